// rp_host_bench.cpp -- CPU-only timing of the root-parallel tree engine (gomokuai_b200/csrc/host/root_parallel.cpp)
// against a STUB of the C-ABI: rollouts are replaced by hash-derived outcomes after a modelled GPU latency, so the
// host-side selection / backup cost can be profiled without a GPU.  Development tool, not part of the product.
//   g++ -O2 -std=c++17 -I include scripts/host_bw/rp_host_bench.cpp gomokuai_b200/csrc/host/root_parallel.cpp \
//       gomokuai_b200/csrc/host/mcts.cpp -lpthread -o /tmp/rp_host_bench && /tmp/rp_host_bench 2048 245 8 100 [groups [reps]]
// tests/test_root_parallel_host.py builds it (also with -fsanitize=thread) to check the scheduler without a GPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../../include/gomoku_b200.h"
#include "../../gomokuai_b200/csrc/host/root_parallel.h"

static int g_latency_us = 100;
static void spin_us(int us) {
    const auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() < us) {}
}
static void fake(const uint32_t* boards, int n, int rollouts, uint64_t key, uint32_t ctr, int base, int32_t* wdb) {
    for (int i = 0; i < n; ++i) {
        uint64_t h = key * 0x9E3779B97F4A7C15ull + ctr * 0xD2511F53ull + uint64_t(base + i) * 0xCD9E8D57ull;
        for (int w = 0; w < 16; ++w) h = (h ^ boards[i * 16 + w]) * 0x100000001B3ull;
        const int b = int((h >> 20) % uint64_t(rollouts + 1));
        wdb[i * 3 + 2] = b; wdb[i * 3 + 0] = rollouts - b; wdb[i * 3 + 1] = 0;
    }
}
extern "C" {
const char* gk_last_error(void) { return "stub"; }
gk_status gk_init(int) { return GK_OK; }
gk_status gk_device_info(int* d, int*, int*, int*) { if (d) *d = 0; return GK_OK; }
gk_status gk_host_alloc(void** out, size_t bytes) { *out = std::malloc(bytes); return GK_OK; }
gk_status gk_host_free(void* p) { std::free(p); return GK_OK; }
gk_status gk_table_default(gk_table**) { return GK_ERR_NO_DEVICE; }
gk_status gk_hybrid_simulate_batch_host(const gk_table*, const uint32_t*, int, float*, float*, int8_t*) { return GK_ERR_NO_DEVICE; }
gk_status gk_rollout_trace_host(const uint32_t*, int, uint64_t, uint32_t, int, int8_t*, uint8_t*, uint8_t*) { return GK_ERR_NO_DEVICE; }
gk_status gk_rollout_batch_host(const uint32_t* b, int n, int r, uint64_t key, uint32_t ctr, int base, int32_t* wdb) {
    spin_us(g_latency_us);
    fake(b, n, r, key, ctr, base, wdb);
    return GK_OK;
}
// asynchronous pair: the stub "GPU" finishes g_latency_us after the submit
struct Slot { std::chrono::steady_clock::time_point ready; } g_slots[8];
gk_status gk_rollout_submit_host(int slot, const uint32_t* b, int n, int r, uint64_t key, uint32_t ctr, int base, int32_t* wdb) {
    fake(b, n, r, key, ctr, base, wdb);
    g_slots[slot].ready = std::chrono::steady_clock::now() + std::chrono::microseconds(g_latency_us);
    return GK_OK;
}
gk_status gk_rollout_wait(int slot) {
    while (std::chrono::steady_clock::now() < g_slots[slot].ready) {}
    return GK_OK;
}
}

int main(int argc, char** argv) {
    using namespace gomoku;
    RootParallelConfig cfg;
    cfg.trees = argc > 1 ? std::atoi(argv[1]) : 2048;
    const int per_tree = argc > 2 ? std::atoi(argv[2]) : 245;
    cfg.threads = argc > 3 ? std::atoi(argv[3]) : 8;
    g_latency_us = argc > 4 ? std::atoi(argv[4]) : 100;
    cfg.groups = argc > 5 ? std::atoi(argv[5]) : 0;
    const int reps = argc > 6 ? std::atoi(argv[6]) : 3;
    cfg.seed = 11;
    Board b;
    for (int c : { 112, 113, 97, 98 }) b.applyMove(Position(c));
    RootParallelSearch s(cfg);
    for (int rep = 0; rep < reps; ++rep) {
        s.run(b, per_tree);
        long long visits = 0, chk = 0;
        for (int c = 0; c < BOARD_SIZE; ++c) { visits += s.stats()[c]; chk = chk * 31 + s.stats()[c]; }
        std::printf("run %d: %.1f ms total, %.1f ms in the stub GPU, %lld playouts -> %.2f M playouts/s, nodes %lld, best %d, chk %llx\n", rep,
                    s.seconds_total * 1e3, s.seconds_gpu * 1e3, visits, visits / s.seconds_total * 1e-6, (long long)s.nodes,
                    int(RootParallelSearch::bestMove(s.stats())), (unsigned long long)chk);
    }
    return 0;
}
