// gk_capi.cu -- the C-ABI of include/gomoku_b200.h: argument checking, table upload,
// stream plumbing and the chunked host<->device pipelines.  No compute happens here.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gomoku_b200.h"
#include "gk_kernels.h"
#include "gk_table.h"

struct gk_table {
    gk::HostTable host;
    uint32_t* d_trans = nullptr;       // host-format words (scan kernel)
    uint16_t* d_next16 = nullptr;      // device-format automaton (eval kernel)
    uint32_t* d_erec = nullptr;
    gk::PatRec* d_patrec = nullptr;
    uint16_t* d_tape_src = nullptr;
    uint16_t* d_tape_info = nullptr;
    int16_t* d_flush = nullptr;
    bool is_default = false;
};

namespace {

thread_local std::string t_error;
std::mutex g_mutex;
int g_device = -1, g_sm_count = 0, g_cc_major = 0, g_cc_minor = 0;
gk_table* g_default_table = nullptr;

gk_status fail(gk_status code, const std::string& msg) { t_error = msg; return code; }
gk_status cuda_fail(cudaError_t e, const char* what) {
    return fail(GK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define GK_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)
// inside a pipelined host call: work already enqueued on the caller's buffers finishes before the error returns
#define GK_CUDA_PIPE(call)                                                      \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) return drain_pipes(cuda_fail(e_, #call));        \
    } while (0)

// The CUDA current device is per THREAD: every entry point rebinds its calling thread to the library's device (a
// Python worker thread, the root-parallel driver's team, ...), so that a process bound to GPU k never touches GPU 0.
thread_local int t_bound_device = -1;
gk_status require_device() {
    if (g_device < 0) return fail(GK_ERR_NOT_INIT, "gk_init() has not been called");
    if (t_bound_device != g_device) {
        GK_CUDA(cudaSetDevice(g_device));
        t_bound_device = g_device;
    }
    return GK_OK;
}

template <class T>
gk_status to_device(T** dst, const std::vector<T>& src) {
    GK_CUDA(cudaMalloc(dst, src.size() * sizeof(T)));
    GK_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GK_OK;
}

gk_status upload(gk_table* t) {
    const gk::HostTable& h = t->host;
    if (gk_status s = to_device(&t->d_next16, h.dev_next)) return s;
    if (gk_status s = to_device(&t->d_erec, h.dev_erec)) return s;
    if (gk_status s = to_device(&t->d_patrec, h.patrec)) return s;
    if (gk_status s = to_device(&t->d_tape_src, h.tape_src)) return s;
    if (gk_status s = to_device(&t->d_tape_info, h.tape_info)) return s;
    if (gk_status s = to_device(&t->d_flush, h.flush)) return s;
    return to_device(&t->d_trans, h.trans);    // last: its presence marks the table as uploaded
}

void free_device_copies(gk_table* t) {
    cudaFree(t->d_trans); cudaFree(t->d_next16); cudaFree(t->d_erec); cudaFree(t->d_patrec); cudaFree(t->d_tape_src);
    cudaFree(t->d_tape_info); cudaFree(t->d_flush);
    t->d_trans = nullptr; t->d_next16 = nullptr; t->d_erec = nullptr; t->d_patrec = nullptr; t->d_tape_src = nullptr;
    t->d_tape_info = nullptr; t->d_flush = nullptr;
}

// device copies are created lazily so that tables can be compiled and inspected without a GPU
gk_status ensure_uploaded(const gk_table* ct) {
    gk_table* t = const_cast<gk_table*>(ct);
    std::lock_guard<std::mutex> lock(g_mutex);
    if (t->d_trans) return GK_OK;
    return upload(t);
}

gk::EvalArgs eval_args(const gk_table* t, const uint32_t* boards, long long n, int32_t* scores, uint16_t* pat,
                       uint16_t* cmp, int8_t* winner) {
    gk::EvalArgs a{};
    a.next16 = t->d_next16; a.erec = t->d_erec; a.n_states = t->host.n_states; a.n_clones = t->host.n_clones;
    a.patrec = t->d_patrec; a.n_patterns = (int)t->host.patrec.size();
    a.tape_src = t->d_tape_src; a.tape_info = t->d_tape_info; a.tape_steps = t->host.tape_steps;
    a.root_off = t->host.root_off; a.start_off = t->host.start_off; a.list_cap = t->host.list_cap; a.trail_pad = t->host.trail_pad;
    a.boards = boards; a.n = n;
    a.scores = scores; a.pat_totals = pat; a.cmp_totals = cmp; a.winner = winner;
    return a;
}

// ---- Philox4x32-10 on the host (synthetic position generator only) -----------------------------
void philox_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        c1 = uint32_t(p1); c3 = uint32_t(p0); c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

bool makes_five(const uint8_t* cells, int c, int colour) {
    const int x0 = c % 15, y0 = c / 15;
    static const int DX[4] = { 1, 0, 1, 1 }, DY[4] = { 0, 1, 1, -1 };
    for (int d = 0; d < 4; ++d) {
        int run = 1;
        for (int s = -1; s <= 1; s += 2)
            for (int i = 1; i < 5; ++i) {
                const int x = x0 + s * i * DX[d], y = y0 + s * i * DY[d];
                if (x < 0 || x >= 15 || y < 0 || y >= 15 || cells[y * 15 + x] != colour) break;
                ++run;
            }
        if (run >= 5) return true;
    }
    return false;
}

void synth_one(int64_t index, uint32_t* board, int16_t* moves, int* n_moves) {
    const uint32_t key[2] = { 0x4F4B5531u, 0x00474F4Du };          // 0x474F4D4F4B5531 "GOMOKU1"
    const uint32_t ilo = uint32_t(index), ihi = uint32_t(uint64_t(index) >> 32);
    uint32_t w[4];
    const uint32_t hctr[4] = { 0xffffffffu, 0u, ilo, ihi };
    philox_host(hctr, key, w);
    const int target = 16 + int(w[0] % 81u);
    uint8_t cells[225] = { 0 };
    int16_t empty[225];
    for (int i = 0; i < 225; ++i) empty[i] = (int16_t)i;
    int n_empty = 225, placed = 0;
    uint32_t draw = 0, block = 0xffffffffu;
    while (placed < target) {
        const int colour = (placed & 1) ? 2 : 1;
        int slot = -1;
        for (int tries = 0; tries < 64 && slot < 0; ++tries, ++draw) {
            if ((draw >> 2) != block) {
                block = draw >> 2;
                const uint32_t ctr[4] = { block, 1u, ilo, ihi };
                philox_host(ctr, key, w);
            }
            const int k = int((uint64_t(w[draw & 3u]) * uint64_t(n_empty)) >> 32);
            if (!makes_five(cells, empty[k], colour)) slot = k;    // a stone that completes five-or-more is redrawn
        }
        if (slot < 0) break;                                        // 64 redraws all completed a five: stop short
        const int c = empty[slot];
        cells[c] = (uint8_t)colour;
        for (int i = slot; i + 1 < n_empty; ++i) empty[i] = empty[i + 1];   // keep increasing cell order
        --n_empty;
        if (moves) moves[placed] = (int16_t)c;
        ++placed;
    }
    for (int i = 0; i < 16; ++i) board[i] = 0;
    for (int c = 0; c < 225; ++c) board[c >> 4] |= uint32_t(cells[c]) << ((c & 15) * 2);
    *n_moves = placed;
}

// ---- pipelined host entry points -----------------------------------------------------------------
struct Pipe {
    cudaStream_t stream = nullptr;
    uint32_t* d_boards = nullptr; int32_t* d_scores = nullptr; uint16_t* d_pat = nullptr; uint16_t* d_cmp = nullptr;
    int8_t* d_win = nullptr; int32_t* d_wdb = nullptr;
    int cap = 0;            // positions the evaluation buffers hold
    int cap_rollout = 0;    // positions d_boards / d_wdb hold (>= cap)
};
constexpr int kPipes = 3;
Pipe g_pipes[kPipes];

void pipe_release(Pipe& p) {                         // buffers only; the stream stays
    if (p.stream) cudaStreamSynchronize(p.stream);
    cudaFree(p.d_boards); cudaFree(p.d_scores); cudaFree(p.d_pat); cudaFree(p.d_cmp); cudaFree(p.d_win); cudaFree(p.d_wdb);
    p.d_boards = nullptr; p.d_scores = nullptr; p.d_pat = nullptr; p.d_cmp = nullptr; p.d_win = nullptr; p.d_wdb = nullptr;
    p.cap = p.cap_rollout = 0;
}

// buffers of the evaluation pipes (3.7 KB per position)
gk_status pipe_reserve(Pipe& p, int chunk) {
    if (!p.stream) GK_CUDA(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    if (p.cap >= chunk) return GK_OK;
    pipe_release(p);
    cudaError_t e = cudaMalloc(&p.d_boards, size_t(chunk) * 64);
    if (e == cudaSuccess) e = cudaMalloc(&p.d_scores, size_t(chunk) * 3600);
    if (e == cudaSuccess) e = cudaMalloc(&p.d_pat, size_t(chunk) * 32);
    if (e == cudaSuccess) e = cudaMalloc(&p.d_cmp, size_t(chunk) * 12);
    if (e == cudaSuccess) e = cudaMalloc(&p.d_win, size_t(chunk));
    if (e == cudaSuccess) e = cudaMalloc(&p.d_wdb, size_t(chunk) * 12);
    if (e != cudaSuccess) { pipe_release(p); return cuda_fail(e, "cudaMalloc (evaluation pipe)"); }
    p.cap = p.cap_rollout = chunk;
    return GK_OK;
}

// the rollout host path needs boards and counts only (76 B per position, not the 3.7 KB of an evaluation pipe)
gk_status pipe_reserve_rollout(Pipe& p, int n) {
    if (!p.stream) GK_CUDA(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    if (p.cap_rollout >= n) return GK_OK;
    pipe_release(p);
    cudaError_t e = cudaMalloc(&p.d_boards, size_t(n) * 64);
    if (e == cudaSuccess) e = cudaMalloc(&p.d_wdb, size_t(n) * 12);
    if (e != cudaSuccess) { pipe_release(p); return cuda_fail(e, "cudaMalloc (rollout pipe)"); }
    p.cap_rollout = n;
    return GK_OK;
}

// an error in the middle of a pipelined call: let the work already enqueued on the caller's buffers finish first
gk_status drain_pipes(gk_status s) {
    for (Pipe& p : g_pipes) if (p.stream) cudaStreamSynchronize(p.stream);
    return s;
}

}  // namespace

extern "C" {

const char* gk_version(void) { return "gomoku_b200 0.1 (sm_100a)"; }
const char* gk_last_error(void) { return t_error.c_str(); }

gk_status gk_init(int device) {
    std::lock_guard<std::mutex> lock(g_mutex);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(GK_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
    if (device < 0 || device >= count) return fail(GK_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop{};
    GK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GK_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100; this library carries sm_100a code only");
    GK_CUDA(cudaSetDevice(device));
    t_bound_device = device;
    g_device = device; g_sm_count = prop.multiProcessorCount; g_cc_major = prop.major; g_cc_minor = prop.minor;
    return GK_OK;
}

static void release_late_resources();       // buffers owned by the entry points further down (defined at the end of the file)

gk_status gk_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    release_late_resources();
    for (Pipe& p : g_pipes) {
        pipe_release(p);
        if (p.stream) cudaStreamDestroy(p.stream);
        p = Pipe{};
    }
    if (g_default_table) free_device_copies(g_default_table);        // re-uploaded lazily on the device of the next gk_init
    g_device = -1;
    return GK_OK;
}

gk_status gk_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor) {
    if (gk_status s = require_device()) return s;
    if (device) *device = g_device;
    if (sm_count) *sm_count = g_sm_count;
    if (cc_major) *cc_major = g_cc_major;
    if (cc_minor) *cc_minor = g_cc_minor;
    return GK_OK;
}

gk_status gk_measure_issue_peak(int mode, double* warp_inst_per_s) {
    if (gk_status s = require_device()) return s;
    if (mode < 0 || mode > 2 || !warp_inst_per_s) return fail(GK_ERR_INVALID, "mode must be 0, 1 or 2");
    GK_CUDA(gk::measure_issue_peak(mode, g_sm_count, 4096, warp_inst_per_s, nullptr));
    return GK_OK;
}

// ---- tables ------------------------------------------------------------------------------------
gk_status gk_table_default(gk_table** out) {
    if (!out) return fail(GK_ERR_INVALID, "out is null");
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_default_table) {
        gk_table* t = new gk_table;
        if (!gk::compile_table(gk::default_protos(), t->host)) {
            const std::string msg = t->host.error;
            delete t;
            return fail(GK_ERR_TABLE, msg);
        }
        t->is_default = true;
        g_default_table = t;
    }
    *out = g_default_table;
    return GK_OK;
}

gk_status gk_table_build(const char* const* protos, const int* types, const int* scores, int n, gk_table** out) {
    if (!protos || !types || !scores || !out || n <= 0) return fail(GK_ERR_INVALID, "bad arguments");
    std::vector<gk::Proto> v;
    for (int i = 0; i < n; ++i) {
        if (!protos[i]) return fail(GK_ERR_INVALID, "null prototype");
        v.push_back({ protos[i], types[i], scores[i] });
    }
    gk_table* t = new gk_table;
    if (!gk::compile_table(v, t->host)) {
        const std::string msg = t->host.error;
        delete t;
        return fail(GK_ERR_TABLE, msg);
    }
    *out = t;
    return GK_OK;
}

gk_status gk_table_free(gk_table* t) {
    if (!t || t->is_default) return GK_OK;
    free_device_copies(t);
    delete t;
    return GK_OK;
}

gk_status gk_table_info(const gk_table* t, int* n_states, int* n_patterns, int* trail_pad, int* max_steps) {
    if (!t) return fail(GK_ERR_INVALID, "table is null");
    if (n_states) *n_states = t->host.n_states;
    if (n_patterns) *n_patterns = (int)t->host.patterns.size();
    if (trail_pad) *trail_pad = t->host.trail_pad;
    if (max_steps) *max_steps = t->host.tape_steps;
    return GK_OK;
}

gk_status gk_table_pattern(const gk_table* t, int id, char str8[8], int* favour, int* type, int* score) {
    if (!t || id < 0 || id >= (int)t->host.patterns.size()) return fail(GK_ERR_INVALID, "pattern id out of range");
    const gk::PatternInfo& p = t->host.patterns[id];
    if (str8) { std::memset(str8, 0, 8); std::memcpy(str8, p.str.data(), p.str.size()); }
    if (favour) *favour = p.favour;
    if (type) *type = p.type;
    if (score) *score = p.score;
    return GK_OK;
}

gk_status gk_table_entries(const gk_table* t, uint32_t* h_entries, int capacity) {
    if (!t || !h_entries || capacity < (int)t->host.trans.size()) return fail(GK_ERR_INVALID, "buffer too small");
    std::memcpy(h_entries, t->host.trans.data(), t->host.trans.size() * sizeof(uint32_t));
    return GK_OK;
}

gk_status gk_table_flush(const gk_table* t, int16_t* h_flush, int capacity) {
    if (!t || !h_flush || capacity < (int)t->host.flush.size()) return fail(GK_ERR_INVALID, "buffer too small");
    std::memcpy(h_flush, t->host.flush.data(), t->host.flush.size() * sizeof(int16_t));
    return GK_OK;
}

gk_status gk_table_device_format(const gk_table* t, int info[6], uint16_t* h_next, int next_capacity, uint32_t* h_erec,
                                 int erec_capacity) {
    if (!t || !info) return fail(GK_ERR_INVALID, "bad arguments");
    const gk::HostTable& h = t->host;
    info[0] = h.n_clones + h.n_states; info[1] = h.n_clones; info[2] = h.root_off; info[3] = h.start_off;
    info[4] = h.list_cap; info[5] = h.tape_steps;
    if (h_next) {
        if (next_capacity < (int)h.dev_next.size()) return fail(GK_ERR_INVALID, "buffer too small");
        std::memcpy(h_next, h.dev_next.data(), h.dev_next.size() * sizeof(uint16_t));
    }
    if (h_erec) {
        if (erec_capacity < (int)h.dev_erec.size()) return fail(GK_ERR_INVALID, "buffer too small");
        std::memcpy(h_erec, h.dev_erec.data(), h.dev_erec.size() * sizeof(uint32_t));
    }
    return GK_OK;
}

// ---- scan ----------------------------------------------------------------------------------------
gk_status gk_scan_batch(const gk_table* t, const uint8_t* d_codes, const int64_t* d_starts, int n_strings,
                        int max_per_string, int32_t* d_pids, int32_t* d_offsets, int32_t* d_counts, void* stream) {
    if (gk_status s = require_device()) return s;
    if (!t || !d_codes || !d_starts || !d_pids || !d_offsets || !d_counts || n_strings < 0 || max_per_string <= 0)
        return fail(GK_ERR_INVALID, "bad arguments");
    if (gk_status s = ensure_uploaded(t)) return s;
    gk::ScanArgs a{ t->d_trans, t->d_flush, d_codes, reinterpret_cast<const long long*>(d_starts), n_strings,
                    max_per_string, d_pids, d_offsets, d_counts };
    GK_CUDA(gk::launch_scan(a, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

// ---- eval ----------------------------------------------------------------------------------------
gk_status gk_eval_batch(const gk_table* t, const uint32_t* d_boards, int n, int32_t* d_scores, uint16_t* d_pat_totals,
                        uint16_t* d_cmp_totals, int8_t* d_winner, void* stream) {
    if (gk_status s = require_device()) return s;
    if (!t || n < 0 || (n > 0 && !d_boards)) return fail(GK_ERR_INVALID, "bad arguments");
    if (reinterpret_cast<uintptr_t>(d_scores) & 15u) return fail(GK_ERR_INVALID, "d_scores must be 16-byte aligned");
    if (gk_status s = ensure_uploaded(t)) return s;
    const gk::EvalArgs a = eval_args(t, d_boards, n, d_scores, d_pat_totals, d_cmp_totals, d_winner);
    GK_CUDA(gk::launch_eval(a, g_sm_count, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

gk_status gk_eval_batch_host(const gk_table* t, const uint32_t* h_boards, int n, int32_t* h_scores, uint16_t* h_pat,
                             uint16_t* h_cmp, int8_t* h_win) {
    if (gk_status s = require_device()) return s;
    if (!t || n < 0 || (n > 0 && !h_boards)) return fail(GK_ERR_INVALID, "bad arguments");
    if (gk_status s = ensure_uploaded(t)) return s;
    // three chunks in flight: copy-in of chunk k+1 and copy-out of chunk k-1 overlap the kernel of chunk k
    const int chunk = std::min(n, 16384);
    if (chunk == 0) return GK_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    for (Pipe& p : g_pipes) if (gk_status s = pipe_reserve(p, chunk)) return s;
    int k = 0;
    for (int at = 0; at < n; at += chunk, ++k) {
        Pipe& p = g_pipes[k % kPipes];
        const int m = std::min(chunk, n - at);
        GK_CUDA_PIPE(cudaMemcpyAsync(p.d_boards, h_boards + size_t(at) * 16, size_t(m) * 64, cudaMemcpyHostToDevice, p.stream));
        const gk::EvalArgs a = eval_args(t, p.d_boards, m, h_scores ? p.d_scores : nullptr, h_pat ? p.d_pat : nullptr,
                                         h_cmp ? p.d_cmp : nullptr, h_win ? p.d_win : nullptr);
        GK_CUDA_PIPE(gk::launch_eval(a, g_sm_count, p.stream));
        if (h_scores) GK_CUDA_PIPE(cudaMemcpyAsync(h_scores + size_t(at) * 900, p.d_scores, size_t(m) * 3600, cudaMemcpyDeviceToHost, p.stream));
        if (h_pat) GK_CUDA_PIPE(cudaMemcpyAsync(h_pat + size_t(at) * 16, p.d_pat, size_t(m) * 32, cudaMemcpyDeviceToHost, p.stream));
        if (h_cmp) GK_CUDA_PIPE(cudaMemcpyAsync(h_cmp + size_t(at) * 6, p.d_cmp, size_t(m) * 12, cudaMemcpyDeviceToHost, p.stream));
        if (h_win) GK_CUDA_PIPE(cudaMemcpyAsync(h_win + at, p.d_win, size_t(m), cudaMemcpyDeviceToHost, p.stream));
    }
    for (Pipe& p : g_pipes) GK_CUDA(cudaStreamSynchronize(p.stream));
    return GK_OK;
}

gk_status gk_eval_policy_batch(const gk_table* t, const uint32_t* d_boards, int n, float* d_probs, float* d_value,
                               int32_t* d_scores, uint16_t* d_pat_totals, uint16_t* d_cmp_totals, int8_t* d_winner,
                               void* stream) {
    if (gk_status s = require_device()) return s;
    if (!t || n < 0 || (n > 0 && !d_boards)) return fail(GK_ERR_INVALID, "bad arguments");
    if (reinterpret_cast<uintptr_t>(d_scores) & 15u) return fail(GK_ERR_INVALID, "d_scores must be 16-byte aligned");
    if (gk_status s = ensure_uploaded(t)) return s;
    gk::EvalArgs a = eval_args(t, d_boards, n, d_scores, d_pat_totals, d_cmp_totals, d_winner);
    a.probs = d_probs; a.value = d_value;
    GK_CUDA(gk::launch_eval(a, g_sm_count, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

gk_status gk_hybrid_simulate_batch(const gk_table* t, const uint32_t* d_boards, int n, float* d_probs, float* d_value,
                                   int8_t* d_winner, uint32_t* d_dflags, void* stream) {
    if (gk_status s = require_device()) return s;
    if (!t || n < 0 || (n > 0 && (!d_boards || !d_probs))) return fail(GK_ERR_INVALID, "bad arguments");
    if (gk_status s = ensure_uploaded(t)) return s;
    gk::EvalArgs a = eval_args(t, d_boards, n, nullptr, nullptr, nullptr, d_winner);
    a.probs = d_probs; a.value = d_value; a.decisive = 1; a.dflags = d_dflags;
    GK_CUDA(gk::launch_eval(a, g_sm_count, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

static gk_status policy_host(const gk_table* t, const uint32_t* h_boards, int n, float* h_probs, float* h_value,
                             int8_t* h_win, int decisive);

gk_status gk_hybrid_simulate_batch_host(const gk_table* t, const uint32_t* h_boards, int n, float* h_probs, float* h_value,
                                        int8_t* h_win) {
    return policy_host(t, h_boards, n, h_probs, h_value, h_win, 1);
}

gk_status gk_eval_policy_batch_host(const gk_table* t, const uint32_t* h_boards, int n, float* h_probs, float* h_value,
                                    int8_t* h_win) {
    return policy_host(t, h_boards, n, h_probs, h_value, h_win, 0);
}

namespace {
constexpr int kSmallPolicyBatch = 64;
unsigned char* g_policy_stage = nullptr;     // page-locked: kSmallPolicyBatch x (64 B board + 905 B of results, padded)
}  // namespace

static gk_status policy_host(const gk_table* t, const uint32_t* h_boards, int n, float* h_probs, float* h_value,
                             int8_t* h_win, int decisive) {
    if (gk_status s = require_device()) return s;
    if (!t || n < 0 || (n > 0 && !h_boards)) return fail(GK_ERR_INVALID, "bad arguments");
    if (gk_status s = ensure_uploaded(t)) return s;
    const int chunk = std::min(n, 16384);
    if (chunk == 0) return GK_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    for (Pipe& p : g_pipes) if (gk_status s = pipe_reserve(p, chunk)) return s;
    if (n <= kSmallPolicyBatch) {
        // one MCTS leaf at a time (TraditionalPolicy::hybridSimulate): pure call latency.  The board is read by the kernel
        // from page-locked staging, probs | value | winner come back in ONE copy, one synchronisation.
        if (!g_policy_stage) GK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_policy_stage), size_t(kSmallPolicyBatch) * (64 + 912), cudaHostAllocDefault));
        uint32_t* st_boards = reinterpret_cast<uint32_t*>(g_policy_stage);
        unsigned char* st_out = g_policy_stage + size_t(kSmallPolicyBatch) * 64;
        std::memcpy(st_boards, h_boards, size_t(n) * 64);
        Pipe& p = g_pipes[0];
        float* d_probs = reinterpret_cast<float*>(p.d_scores);
        float* d_value = d_probs + size_t(n) * 225;
        int8_t* d_winner = reinterpret_cast<int8_t*>(d_value + n);
        gk::EvalArgs a = eval_args(t, st_boards, n, nullptr, nullptr, nullptr, d_winner);
        a.probs = d_probs; a.value = d_value; a.decisive = decisive;
        GK_CUDA(gk::launch_eval(a, g_sm_count, p.stream));
        GK_CUDA(cudaMemcpyAsync(st_out, d_probs, size_t(n) * 905, cudaMemcpyDeviceToHost, p.stream));
        GK_CUDA(cudaStreamSynchronize(p.stream));
        if (h_probs) std::memcpy(h_probs, st_out, size_t(n) * 900);
        if (h_value) std::memcpy(h_value, st_out + size_t(n) * 900, size_t(n) * 4);
        if (h_win) std::memcpy(h_win, st_out + size_t(n) * 904, size_t(n));
        return GK_OK;
    }
    int k = 0;
    for (int at = 0; at < n; at += chunk, ++k) {                      // the score buffer of the pipe doubles as probs + value
        Pipe& p = g_pipes[k % kPipes];
        const int m = std::min(chunk, n - at);
        float* d_probs = reinterpret_cast<float*>(p.d_scores);
        float* d_value = d_probs + size_t(chunk) * 225;
        GK_CUDA_PIPE(cudaMemcpyAsync(p.d_boards, h_boards + size_t(at) * 16, size_t(m) * 64, cudaMemcpyHostToDevice, p.stream));
        gk::EvalArgs a = eval_args(t, p.d_boards, m, nullptr, nullptr, nullptr, h_win ? p.d_win : nullptr);
        a.probs = h_probs ? d_probs : nullptr; a.value = h_value ? d_value : nullptr; a.decisive = decisive;
        if (!a.probs && !a.value) a.value = d_value;
        GK_CUDA_PIPE(gk::launch_eval(a, g_sm_count, p.stream));
        if (h_probs) GK_CUDA_PIPE(cudaMemcpyAsync(h_probs + size_t(at) * 225, d_probs, size_t(m) * 900, cudaMemcpyDeviceToHost, p.stream));
        if (h_value) GK_CUDA_PIPE(cudaMemcpyAsync(h_value + at, d_value, size_t(m) * 4, cudaMemcpyDeviceToHost, p.stream));
        if (h_win) GK_CUDA_PIPE(cudaMemcpyAsync(h_win + at, p.d_win, size_t(m), cudaMemcpyDeviceToHost, p.stream));
    }
    for (Pipe& p : g_pipes) GK_CUDA(cudaStreamSynchronize(p.stream));
    return GK_OK;
}

gk_status gk_guided_rollout_batch(const gk_table* t, const uint32_t* d_boards, int n, int mode, uint64_t philox_key,
                                  uint32_t ctr_hi, int game_base, int max_moves, int8_t* d_winner, int16_t* d_length,
                                  int16_t* d_moves, uint32_t* d_final_boards, void* stream) {
    return gk_guided_rollout_queue(t, d_boards, n, 0, mode, philox_key, ctr_hi, game_base, max_moves, d_winner, d_length, d_moves,
                                   d_final_boards, stream);
}

gk_status gk_guided_rollout_queue(const gk_table* t, const uint32_t* d_boards, int n, int max_in_flight, int mode, uint64_t philox_key,
                                  uint32_t ctr_hi, int game_base, int max_moves, int8_t* d_winner, int16_t* d_length,
                                  int16_t* d_moves, uint32_t* d_final_boards, void* stream) {
    if (gk_status s = require_device()) return s;
    const int full_rescan = (mode & GK_GUIDED_FULL_RESCAN) ? 1 : 0, single_warp = (mode & GK_GUIDED_SINGLE_WARP) ? 1 : 0;
    mode &= ~(GK_GUIDED_FULL_RESCAN | GK_GUIDED_SINGLE_WARP);
    if (!t || n < 0 || (n > 0 && !d_boards) || (mode != 1 && mode != 2) || max_moves < 0 || max_moves > GK_CELLS)
        return fail(GK_ERR_INVALID, "bad arguments");
    if (gk_status s = ensure_uploaded(t)) return s;
    gk::EvalArgs a = eval_args(t, d_boards, n, nullptr, nullptr, nullptr, nullptr);
    a.g_mode = mode; a.g_full_rescan = full_rescan; a.g_single_warp = single_warp; a.g_in_flight = max_in_flight > 0 ? max_in_flight : 0; a.g_key_lo = uint32_t(philox_key); a.g_key_hi = uint32_t(philox_key >> 32); a.g_ctr_hi = ctr_hi;
    a.g_game_base = game_base; a.g_max_moves = max_moves;
    a.g_winner = d_winner; a.g_length = d_length; a.g_moves = d_moves; a.g_final = d_final_boards;
    GK_CUDA(gk::launch_eval(a, g_sm_count, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

// ---- rollouts --------------------------------------------------------------------------------------
// Slot-image scratch of the rollout kernels: one grow-only buffer per stream, so launches in flight on different
// streams never share it.  Growing frees the old buffer with cudaFree, which waits for the device to drain.
namespace {
struct RolloutScratch { uint32_t* images = nullptr; size_t cap = 0; };
std::map<cudaStream_t, RolloutScratch> g_rollout_scratch;
std::mutex g_scratch_mutex;

gk_status rollout_scratch(cudaStream_t stream, int n, uint32_t** out) {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    RolloutScratch& sc = g_rollout_scratch[stream];
    const size_t need = gk::rollout_scratch_bytes(n);
    if (need > sc.cap) {
        if (sc.images) GK_CUDA(cudaFree(sc.images));
        sc.images = nullptr; sc.cap = 0;
        const size_t want = std::max(need, size_t(64) << 10);
        GK_CUDA(cudaMalloc(&sc.images, want));
        sc.cap = want;
    }
    *out = sc.images;
    return GK_OK;
}
}  // namespace

static gk_status rollout_common(const uint32_t* d_boards, int n, int rollouts_per_pos, uint64_t key, uint32_t ctr_hi,
                                int pos_base, const uint8_t* d_r_stream, int stream_stride, int32_t* d_wdb,
                                int8_t* d_winners, uint8_t* d_lengths, cudaStream_t stream) {
    if (n < 0 || rollouts_per_pos <= 0 || (n > 0 && !d_boards)) return fail(GK_ERR_INVALID, "bad arguments");
    if ((unsigned long long)n * (unsigned long long)rollouts_per_pos >= (1ull << 31))
        return fail(GK_ERR_INVALID, "n * rollouts_per_pos must be below 2^31 per call");
    gk::RolloutArgs a{};
    a.boards = d_boards; a.n = n; a.rollouts_per_pos = rollouts_per_pos;
    a.key_lo = uint32_t(key); a.key_hi = uint32_t(key >> 32); a.ctr_hi = ctr_hi; a.pos_base = pos_base;
    a.r_stream = d_r_stream; a.stream_stride = stream_stride;
    a.wdb = d_wdb; a.winners = d_winners; a.lengths = d_lengths;
    if (n == 0) return GK_OK;
    uint32_t* images = nullptr;
    if (gk_status s = rollout_scratch(stream, n, &images)) return s;
    GK_CUDA(gk::launch_rollout(a, g_sm_count, images, stream));
    return GK_OK;
}

gk_status gk_rollout_batch(const uint32_t* d_boards, int n, int rollouts_per_pos, uint64_t philox_key, uint32_t ctr_hi,
                           int pos_base, int32_t* d_wdb, int8_t* d_winners, uint8_t* d_lengths, void* stream) {
    if (gk_status s = require_device()) return s;
    return rollout_common(d_boards, n, rollouts_per_pos, philox_key, ctr_hi, pos_base, nullptr, 0, d_wdb, d_winners,
                          d_lengths, static_cast<cudaStream_t>(stream));
}

gk_status gk_rollout_injected(const uint32_t* d_boards, int n, int rollouts_per_pos, const uint8_t* d_r_stream,
                              int stream_stride, int8_t* d_winners, uint8_t* d_lengths, void* stream) {
    if (gk_status s = require_device()) return s;
    if (!d_r_stream || stream_stride <= 0 || stream_stride > 254) return fail(GK_ERR_INVALID, "bad injected stream");
    return rollout_common(d_boards, n, rollouts_per_pos, 0, 0, 0, d_r_stream, stream_stride, nullptr, d_winners,
                          d_lengths, static_cast<cudaStream_t>(stream));
}

namespace {
// small synchronous batches (one MCTS leaf at a time, config 1) are pure call latency: they go through page-locked
// staging that the kernel reads in place (no copy-in operation) and a true DMA copy-out
constexpr int kSmallBatch = 1024;
constexpr int kFusedBatch = 16;         // up to this many positions: one fused launch (one CTA per position)
uint32_t* g_stage_boards = nullptr;     // page-locked, kSmallBatch x 16 words
int32_t* g_stage_wdb = nullptr;         // page-locked, kSmallBatch x 3

// One-launch rollouts of a small batch: a warp per rollout while the whole batch is one resident wave of warps (the
// latency form: a move costs ~1/2 of the thread-per-rollout chain), else a thread per rollout.  Same results either way.
cudaError_t launch_small(const gk::RolloutArgs& a, int32_t* out, cudaStream_t stream, const uint32_t* h_boards) {
    return gk::rollout_warp_fits(a.n, a.rollouts_per_pos, g_sm_count) ? gk::launch_rollout_warp(a, out, stream, h_boards)
                                                                       : gk::launch_rollout_small(a, out, stream);
}
}  // namespace

gk_status gk_rollout_batch_host(const uint32_t* h_boards, int n, int rollouts_per_pos, uint64_t philox_key,
                                uint32_t ctr_hi, int pos_base, int32_t* h_wdb) {
    if (gk_status s = require_device()) return s;
    if (n < 0 || (n > 0 && (!h_boards || !h_wdb))) return fail(GK_ERR_INVALID, "bad arguments");
    if (n == 0) return GK_OK;
    std::lock_guard<std::mutex> lock(g_mutex);
    Pipe& p = g_pipes[0];
    if (gk_status s = pipe_reserve_rollout(p, std::max(n, 1))) return s;
    if (n <= kSmallBatch) {
        if (!g_stage_boards) {
            GK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_stage_boards), size_t(kSmallBatch) * 64, cudaHostAllocDefault));
            GK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_stage_wdb), size_t(kSmallBatch) * 12, cudaHostAllocDefault));
        }
        std::memcpy(g_stage_boards, h_boards, size_t(n) * 64);
        if (n <= kFusedBatch && rollouts_per_pos > 0 && rollouts_per_pos <= 256) {
            // one leaf of a single-tree search: ONE launch that reads the staged boards and writes the staged counts
            gk::RolloutArgs a{};
            a.boards = g_stage_boards; a.n = n; a.rollouts_per_pos = rollouts_per_pos;
            a.key_lo = uint32_t(philox_key); a.key_hi = uint32_t(philox_key >> 32); a.ctr_hi = ctr_hi; a.pos_base = pos_base;
            // the kernel stores a position's three counts (never negative) as that position's block retires: watching them
            // arrive in the page-locked block saves the stream synchronisation's own latency, ~3 us of a ~27 us call
            volatile int32_t* counts = g_stage_wdb;
            for (int k = 0; k < 3 * n; ++k) counts[k] = INT32_MIN;
            GK_CUDA(launch_small(a, g_stage_wdb, p.stream, g_stage_boards));
            int seen = 0;
            for (int spins = 0; spins < (1 << 22); ++spins) {
                while (seen < 3 * n && counts[seen] != INT32_MIN) ++seen;
                if (seen == 3 * n) break;
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            if (seen < 3 * n) GK_CUDA(cudaStreamSynchronize(p.stream));      // nothing arrived: the stream knows why
            std::memcpy(h_wdb, g_stage_wdb, size_t(n) * 12);
            return GK_OK;
        }
        if (gk_status s = rollout_common(g_stage_boards, n, rollouts_per_pos, philox_key, ctr_hi, pos_base, nullptr, 0, p.d_wdb,
                                         nullptr, nullptr, p.stream))
            return s;
        GK_CUDA(cudaMemcpyAsync(g_stage_wdb, p.d_wdb, size_t(n) * 12, cudaMemcpyDeviceToHost, p.stream));
        GK_CUDA(cudaStreamSynchronize(p.stream));
        std::memcpy(h_wdb, g_stage_wdb, size_t(n) * 12);
        return GK_OK;
    }
    GK_CUDA(cudaMemcpyAsync(p.d_boards, h_boards, size_t(n) * 64, cudaMemcpyHostToDevice, p.stream));
    if (gk_status s = rollout_common(p.d_boards, n, rollouts_per_pos, philox_key, ctr_hi, pos_base, nullptr, 0, p.d_wdb,
                                     nullptr, nullptr, p.stream))
        return s;
    GK_CUDA(cudaMemcpyAsync(h_wdb, p.d_wdb, size_t(n) * 12, cudaMemcpyDeviceToHost, p.stream));
    GK_CUDA(cudaStreamSynchronize(p.stream));
    return GK_OK;
}

namespace {
unsigned char* g_stage_trace = nullptr;   // page-locked: 256 x (225 moves + length + winner)
}  // namespace

gk_status gk_rollout_trace_host(const uint32_t* h_board, int rollouts, uint64_t philox_key, uint32_t ctr_hi, int pos,
                                int8_t* h_winners, uint8_t* h_lengths, uint8_t* h_moves) {
    if (gk_status s = require_device()) return s;
    if (!h_board || rollouts <= 0 || rollouts > 256 || !h_winners || !h_lengths || !h_moves) return fail(GK_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lock(g_mutex);
    Pipe& p = g_pipes[0];
    if (!p.stream) GK_CUDA(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    if (!g_stage_boards) {
        GK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_stage_boards), size_t(kSmallBatch) * 64, cudaHostAllocDefault));
        GK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_stage_wdb), size_t(kSmallBatch) * 12, cudaHostAllocDefault));
    }
    if (!g_stage_trace) GK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_stage_trace), size_t(256) * (GK_CELLS + 2), cudaHostAllocDefault));
    std::memcpy(g_stage_boards, h_board, 64);
    gk::RolloutArgs a{};
    a.boards = g_stage_boards; a.n = 1; a.rollouts_per_pos = rollouts;
    a.key_lo = uint32_t(philox_key); a.key_hi = uint32_t(philox_key >> 32); a.ctr_hi = ctr_hi; a.pos_base = pos;
    a.moves = g_stage_trace;
    a.lengths = g_stage_trace + size_t(256) * GK_CELLS;
    a.winners = reinterpret_cast<int8_t*>(g_stage_trace + size_t(256) * (GK_CELLS + 1));
    GK_CUDA(launch_small(a, nullptr, p.stream, g_stage_boards));
    GK_CUDA(cudaStreamSynchronize(p.stream));
    std::memcpy(h_lengths, a.lengths, size_t(rollouts));
    std::memcpy(h_winners, a.winners, size_t(rollouts));
    for (int r = 0; r < rollouts; ++r) std::memcpy(h_moves + size_t(r) * GK_CELLS, a.moves + size_t(r) * GK_CELLS, h_lengths[r]);
    return GK_OK;
}

// ---- asynchronous host entry points: several small batches in flight (root-parallel search) ----------------
namespace {
constexpr int kAsyncSlots = 16;
constexpr int kFusedAsyncMax = 4096;     // positions per asynchronous batch that still go through the one-launch path
struct AsyncSlot {
    std::mutex mutex;
    cudaStream_t stream = nullptr;
    uint32_t* d_boards = nullptr; int32_t* d_wdb = nullptr; int cap = 0;
    // the caller's page-locked buffers of the last call and their device addresses (a search passes the same two every time)
    const void* seen_boards = nullptr; const uint32_t* dev_boards = nullptr;
    const void* seen_wdb = nullptr; int32_t* dev_wdb = nullptr;
};
AsyncSlot g_async[kAsyncSlots];

// device address of a page-locked, mapped host range containing `p`, or null for pageable memory
void* mapped_address(const void* p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        return attr.devicePointer;
    cudaGetLastError();
    return nullptr;
}
}  // namespace

gk_status gk_rollout_submit_host(int slot, const uint32_t* h_boards, int n, int rollouts_per_pos, uint64_t philox_key,
                                 uint32_t ctr_hi, int pos_base, int32_t* h_wdb) {
    if (gk_status s = require_device()) return s;
    if (slot < 0 || slot >= kAsyncSlots) return fail(GK_ERR_INVALID, "slot out of range");
    if (n < 0 || (n > 0 && (!h_boards || !h_wdb))) return fail(GK_ERR_INVALID, "bad arguments");
    if (n == 0) return GK_OK;
    AsyncSlot& a = g_async[slot];
    std::lock_guard<std::mutex> lock(a.mutex);
    if (!a.stream) GK_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    if (a.cap < n) {
        GK_CUDA(cudaStreamSynchronize(a.stream));
        cudaFree(a.d_boards); cudaFree(a.d_wdb);
        a.d_boards = nullptr; a.d_wdb = nullptr; a.cap = 0;
        const int want = std::max(n, 256);
        GK_CUDA(cudaMalloc(&a.d_boards, size_t(want) * 64));
        GK_CUDA(cudaMalloc(&a.d_wdb, size_t(want) * 12));
        a.cap = want;
    }
    // page-locked boards (gk_host_alloc) are read by the kernel where they lie (unified addressing): one copy less
    // on a path whose cost is launch latency; pageable boards are staged through the slot's device buffer
    if (a.seen_boards != h_boards) { a.dev_boards = static_cast<const uint32_t*>(mapped_address(h_boards)); a.seen_boards = h_boards; }
    if (a.seen_wdb != h_wdb) { a.dev_wdb = static_cast<int32_t*>(mapped_address(h_wdb)); a.seen_wdb = h_wdb; }
    const bool boards_mapped = a.dev_boards != nullptr;
    const uint32_t* boards = boards_mapped ? a.dev_boards : a.d_boards;
    // A leaf batch of a tree search is a few hundred positions x a handful of playouts: pure latency.  With both buffers
    // page-locked the whole round trip is ONE launch -- one CTA per position builds the slot image, plays the playouts and
    // writes the three counts straight into the caller's memory -- instead of image kernel + rollout kernel + copy-out.
    if (boards_mapped && a.dev_wdb && n <= kFusedAsyncMax && rollouts_per_pos <= 256) {
        gk::RolloutArgs ra{};
        ra.boards = boards; ra.n = n; ra.rollouts_per_pos = rollouts_per_pos;
        ra.key_lo = uint32_t(philox_key); ra.key_hi = uint32_t(philox_key >> 32); ra.ctr_hi = ctr_hi; ra.pos_base = pos_base;
        GK_CUDA(launch_small(ra, a.dev_wdb, a.stream, h_boards));
        return GK_OK;
    }
    if (!boards_mapped) GK_CUDA(cudaMemcpyAsync(a.d_boards, h_boards, size_t(n) * 64, cudaMemcpyHostToDevice, a.stream));
    if (gk_status s = rollout_common(boards, n, rollouts_per_pos, philox_key, ctr_hi, pos_base, nullptr, 0, a.d_wdb, nullptr,
                                     nullptr, a.stream))
        return s;
    GK_CUDA(cudaMemcpyAsync(h_wdb, a.d_wdb, size_t(n) * 12, cudaMemcpyDeviceToHost, a.stream));
    return GK_OK;
}

gk_status gk_rollout_wait(int slot) {
    if (gk_status s = require_device()) return s;
    if (slot < 0 || slot >= kAsyncSlots) return fail(GK_ERR_INVALID, "slot out of range");
    AsyncSlot& a = g_async[slot];
    cudaStream_t stream;
    { std::lock_guard<std::mutex> lock(a.mutex); stream = a.stream; }
    if (!stream) return GK_OK;
    GK_CUDA(cudaStreamSynchronize(stream));
    return GK_OK;
}

// ---- feature planes ----------------------------------------------------------------------------------------
gk_status gk_encode_states_batch(const uint32_t* d_boards, const int16_t* d_last_moves, int n, int augment,
                                 uint8_t* d_planes, const float* d_probs, float* d_probs_out, void* stream) {
    if (gk_status s = require_device()) return s;
    if (n < 0 || (n > 0 && (!d_boards || !d_planes)) || (d_probs && !d_probs_out))
        return fail(GK_ERR_INVALID, "bad arguments");
    gk::EncodeArgs a{ d_boards, d_last_moves, n, augment ? 1 : 0, d_planes, d_probs, d_probs_out };
    GK_CUDA(gk::launch_encode(a, g_sm_count, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

gk_status gk_expand_games(const uint32_t* d_boards0, const int16_t* d_moves, const int16_t* d_lengths, const int8_t* d_winners,
                          int n, int max_moves, const int64_t* d_starts, uint32_t* d_out_boards, int16_t* d_out_last_moves,
                          int8_t* d_out_z, void* stream) {
    if (gk_status s = require_device()) return s;
    if (n < 0 || max_moves <= 0 || (n > 0 && (!d_boards0 || !d_moves || !d_lengths || !d_starts || !d_out_boards)))
        return fail(GK_ERR_INVALID, "bad arguments");
    gk::ExpandArgs a{ d_boards0, d_moves, d_lengths, d_winners, n, max_moves, reinterpret_cast<const long long*>(d_starts),
                      d_out_boards, d_out_last_moves, d_out_z };
    GK_CUDA(gk::launch_expand_games(a, static_cast<cudaStream_t>(stream)));
    return GK_OK;
}

// ---- NCCL, resolved at run time ---------------------------------------------------------------------------
namespace {
struct NcclId128 { char b[128]; };                                    // ncclUniqueId: passed by value (nccl.h)
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} g_nccl;
void* g_nccl_comm = nullptr;

gk_status nccl_load() {
    if (g_nccl.handle) return GK_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(GK_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
        return fail(GK_ERR_NCCL, "libnccl.so.2 lacks an expected symbol");
    g_nccl.handle = h;
    return GK_OK;
}
gk_status nccl_fail(int rc, const char* what) { return fail(GK_ERR_NCCL, std::string(what) + ": " + g_nccl.GetErrorString(rc)); }
}  // namespace

gk_status gk_nccl_unique_id(uint8_t id[128]) {
    if (!id) return fail(GK_ERR_INVALID, "id is null");
    if (gk_status s = nccl_load()) return s;
    if (int rc = g_nccl.GetUniqueId(id)) return nccl_fail(rc, "ncclGetUniqueId");
    return GK_OK;
}

gk_status gk_nccl_init(const uint8_t id[128], int world_size, int rank) {
    if (gk_status s = require_device()) return s;
    if (!id || world_size <= 0 || rank < 0 || rank >= world_size) return fail(GK_ERR_INVALID, "bad arguments");
    if (gk_status s = nccl_load()) return s;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_nccl_comm) return fail(GK_ERR_INVALID, "gk_nccl_init called twice");
    NcclId128 uid;
    std::memcpy(uid.b, id, 128);
    if (int rc = g_nccl.CommInitRank(&g_nccl_comm, world_size, uid, rank)) { g_nccl_comm = nullptr; return nccl_fail(rc, "ncclCommInitRank"); }
    return GK_OK;
}

gk_status gk_root_allreduce(void* nccl_comm, int64_t* d_stats, void* stream) {
    if (gk_status s = require_device()) return s;
    if (!d_stats) return fail(GK_ERR_INVALID, "d_stats is null");
    if (gk_status s = nccl_load()) return s;
    void* comm = nccl_comm ? nccl_comm : g_nccl_comm;
    if (!comm) return fail(GK_ERR_NOT_INIT, "no communicator: pass one or call gk_nccl_init");
    constexpr int kNcclInt64 = 4, kNcclSum = 0;                      // ncclDataType_t / ncclRedOp_t (nccl.h)
    if (int rc = g_nccl.AllReduce(d_stats, d_stats, 3 * GK_CELLS, kNcclInt64, kNcclSum, comm, static_cast<cudaStream_t>(stream)))
        return nccl_fail(rc, "ncclAllReduce");
    return GK_OK;
}

gk_status gk_nccl_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_nccl_comm) { g_nccl.CommDestroy(g_nccl_comm); g_nccl_comm = nullptr; }
    return GK_OK;
}

gk_status gk_host_alloc(void** out, size_t bytes) {
    if (gk_status s = require_device()) return s;
    if (!out) return fail(GK_ERR_INVALID, "out is null");
    GK_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return GK_OK;
}

gk_status gk_host_free(void* ptr) {
    if (ptr) {
        for (AsyncSlot& a : g_async) {                  // the slots remember device addresses of page-locked ranges
            std::lock_guard<std::mutex> lock(a.mutex);
            a.seen_boards = a.seen_wdb = nullptr; a.dev_boards = nullptr; a.dev_wdb = nullptr;
        }
        GK_CUDA(cudaFreeHost(ptr));
    }
    return GK_OK;
}

// ---- host utilities ------------------------------------------------------------------------------------
gk_status gk_measure_host_write_bw(int threads, size_t bytes_per_thread, int repeats, double* gb_per_s) {
    if (threads <= 0 || threads > 256 || bytes_per_thread < (size_t(1) << 20) || repeats <= 0 || !gb_per_s)
        return fail(GK_ERR_INVALID, "bad arguments");
    std::vector<unsigned char*> bufs(threads, nullptr);
    for (auto& b : bufs) {
        b = static_cast<unsigned char*>(std::malloc(bytes_per_thread));
        if (!b) { for (auto* q : bufs) std::free(q); return fail(GK_ERR_INVALID, "out of host memory"); }
    }
    std::vector<double> secs(threads, 0.0);
    std::atomic<int> ready{ 0 };
    std::atomic<bool> go{ false };
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w)
        pool.emplace_back([&, w]() {
            std::memset(bufs[w], 1, bytes_per_thread);                   // first touch: the pages exist before the clock starts
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            const auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < repeats; ++r) {
                std::memset(bufs[w], r + 2, bytes_per_thread);           // large memset: glibc streams with non-temporal stores
                asm volatile("" :: "r"(bufs[w]) : "memory");
            }
            secs[w] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
    while (ready.load() < threads) std::this_thread::yield();
    go.store(true, std::memory_order_release);
    for (auto& th : pool) th.join();
    double slowest = 0;
    for (double v : secs) slowest = std::max(slowest, v);
    for (auto* q : bufs) std::free(q);
    *gb_per_s = double(threads) * double(bytes_per_thread) * repeats / slowest / 1e9;
    return GK_OK;
}

gk_status gk_pack_moves(const int16_t* moves, const int64_t* starts, int n, uint32_t* h_boards) {
    if (!moves || !starts || !h_boards || n < 0) return fail(GK_ERR_INVALID, "bad arguments");
    for (int p = 0; p < n; ++p) {
        uint32_t* b = h_boards + size_t(p) * 16;
        for (int i = 0; i < 16; ++i) b[i] = 0;
        int k = 0;
        for (int64_t i = starts[p]; i < starts[p + 1]; ++i, ++k) {
            const int c = moves[i];
            if (c < 0 || c >= 225) return fail(GK_ERR_INVALID, "move out of range");
            if ((b[c >> 4] >> ((c & 15) * 2)) & 3u) return fail(GK_ERR_INVALID, "cell played twice");
            b[c >> 4] |= uint32_t((k & 1) ? 2 : 1) << ((c & 15) * 2);
        }
    }
    return GK_OK;
}

gk_status gk_synth_positions(int64_t first, int n, uint32_t* h_boards, int16_t* h_moves, int64_t* h_starts) {
    if (!h_boards || n < 0 || (h_moves && !h_starts)) return fail(GK_ERR_INVALID, "bad arguments");
    std::vector<int> counts(n);
    const int n_threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int w = 0; w < n_threads; ++w)
        pool.emplace_back([=, &counts]() {
            for (int i = w; i < n; i += n_threads)
                synth_one(first + i, h_boards + size_t(i) * 16, h_moves ? h_moves + size_t(i) * 96 : nullptr, &counts[i]);
        });
    for (std::thread& th : pool) th.join();
    if (h_moves) {                                                   // compact the 96-wide rows into one list
        int64_t at = 0;
        for (int i = 0; i < n; ++i) {
            h_starts[i] = at;
            std::memmove(h_moves + at, h_moves + size_t(i) * 96, size_t(counts[i]) * sizeof(int16_t));
            at += counts[i];
        }
        h_starts[n] = at;
    }
    return GK_OK;
}

}  // extern "C"

// everything gk_shutdown() has to give back besides the pipes: streams, device and page-locked buffers of the entry
// points above, and the library's own NCCL communicator (caller holds g_mutex)
static void release_late_resources() {
    {
        std::lock_guard<std::mutex> lock(g_scratch_mutex);
        for (auto& kv : g_rollout_scratch) cudaFree(kv.second.images);
        g_rollout_scratch.clear();
    }
    for (AsyncSlot& a : g_async) {
        std::lock_guard<std::mutex> lock(a.mutex);
        if (a.stream) { cudaStreamSynchronize(a.stream); cudaStreamDestroy(a.stream); }
        cudaFree(a.d_boards); cudaFree(a.d_wdb);
        a.stream = nullptr; a.d_boards = nullptr; a.d_wdb = nullptr; a.cap = 0;
        a.seen_boards = a.seen_wdb = nullptr; a.dev_boards = nullptr; a.dev_wdb = nullptr;
    }
    cudaFreeHost(g_stage_boards); cudaFreeHost(g_stage_wdb); cudaFreeHost(g_policy_stage); cudaFreeHost(g_stage_trace);
    g_stage_boards = nullptr; g_stage_wdb = nullptr; g_policy_stage = nullptr; g_stage_trace = nullptr;
    if (g_nccl_comm) { g_nccl.CommDestroy(g_nccl_comm); g_nccl_comm = nullptr; }
}
