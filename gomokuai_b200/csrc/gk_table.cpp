// gk_table.cpp -- host-side table compiler.
//
// Input : pattern prototypes in the reference's notation (src/Pattern.cpp:554-596).
// Output: (1) a flat transducer T[state][symbol] -> {next state, <= 2 emissions} whose emission
//             stream is identical to the reference's PatternSearch::generator (src/Pattern.cpp:
//             33-62) on every input string, and (2) the per-lane scan tape of the eval kernel.
//
// Why the compiler goes through the reference's own construction instead of building a
// textbook Aho-Corasick automaton from the pattern set: the reference's automaton is NOT the
// textbook one.  (a) It has no output links; instead it emits on arrival at a terminal node,
// emits again when a FAIL transition lands on a terminal node (before the pending symbol is
// consumed), and collapses runs at four "invariant" states (ACAutomata.cpp:269-272,
// Pattern.cpp:40-45).  (b) Its pattern order comes from an unstable std::sort over keys that
// collide for three pairs of patterns (ACAutomata.cpp:66-90), and its trie is an ordered set
// keyed by (depth, first-pattern-index) only (ACAutomata.h:22-24), so that with the order
// libstdc++ produces, "-x--xx-" (id 218), "-o--oo-" (239) and "--x-xx-" (251) are unreachable
// and their 6-symbol prefixes report ids 217 / 238 / 250 instead.  A drop-in must reproduce
// exactly that, for any prototype list, so stages 1-4 below re-derive the reference's goto /
// fail / output functions step by step (in their own data structures), and stage 5 flattens
// them by simulating the generator once per (state, symbol).
#include "gk_table.h"

#include <algorithm>
#include <cmath>
#include <deque>
#include <map>
#include <numeric>

namespace gk {

const std::vector<Proto>& default_protos() {
    // values of src/Pattern.cpp:554-596; Pattern::Type order DeadOne..Five = 0..8
    static const std::vector<Proto> protos = {
        { "+xxxxx", 8, 9999 },   { "-_oooo_", 7, 9000 },  { "-xoooo_", 6, 2500 },  { "-o_ooo", 6, 3000 },
        { "-oo_oo", 6, 2600 },   { "-~_ooo_~", 5, 3000 }, { "-x^ooo_~", 5, 2900 }, { "-~o_oo~", 5, 2800 },
        { "-~o~oo_~", 4, 1400 }, { "-~oo~o_~", 4, 1200 }, { "-x_o~oo~", 4, 1300 }, { "-x_oo~o~", 4, 1100 },
        { "-xooo__~", 4, 510 },  { "-xoo_o_~", 4, 520 },  { "-xoo__o~", 4, 520 },  { "-xo_oo_~", 4, 530 },
        { "-xo__oo", 4, 530 },   { "-xooo__x", 4, 500 },  { "-xoo_o_x", 4, 500 },  { "-xoo__ox", 4, 500 },
        { "-xo_oo_x", 4, 500 },  { "-x_ooo_x", 4, 500 },  { "-~oo__o~", 4, 750 },  { "-oo__oo", 4, 540 },
        { "-o_o_o", 4, 550 },    { "-~oo__~", 3, 650 },   { "-~_o_o_~", 3, 600 },  { "-x^o_o_^", 3, 550 },
        { "-^o__o^", 3, 550 },   { "-xoo___", 2, 150 },   { "-xo_o__", 2, 160 },   { "-xo__o_", 2, 170 },
        { "-o___o", 2, 180 },    { "-x_oo__x", 2, 120 },  { "-x_o_o_x", 2, 120 },  { "-~o___~", 1, 150 },
        { "-x~_o__^", 1, 140 },  { "-x~__o_^", 1, 150 },  { "-xo___~", 0, 30 },    { "-x_o___x", 0, 40 },
        { "-x__o__x", 0, 50 },
    };
    return protos;
}

namespace {

int code_of(char ch) {   // EncodeCharset, include/Mapping.h:40-48
    switch (ch) {
        case 'x': return 1;
        case 'o': return 2;
        case '?': return 3;
        case '-': case '_': case '^': case '~': return 4;
        default: return 0;
    }
}

// ---- stage 1: augmentation (ACAutomata.cpp:25-64) ------------------------------------------
std::vector<PatternInfo> expand(const std::vector<Proto>& protos) {
    std::vector<PatternInfo> v;
    for (const Proto& p : protos)
        v.push_back({ p.text.substr(1), p.text[0] == '+' ? 1 : -1, p.type, p.score });
    for (size_t i = 0, n = v.size(); i < n; ++i) {            // mirror image, unless a palindrome
        PatternInfo r = v[i];
        std::reverse(r.str.begin(), r.str.end());
        if (r.str != v[i].str) v.push_back(r);
    }
    for (size_t i = 0, n = v.size(); i < n; ++i) {            // colour swap
        PatternInfo f = v[i];
        f.favour = -f.favour;
        for (char& ch : f.str) ch = ch == 'x' ? 'o' : ch == 'o' ? 'x' : ch;
        v.push_back(f);
    }
    for (size_t i = 0, n = v.size(); i < n; ++i) {            // the rival's outermost stones may be the board edge
        const char rival = v[i].favour == 1 ? 'o' : 'x';
        const size_t a = v[i].str.find_first_of(rival), b = v[i].str.find_last_of(rival);
        if (a == std::string::npos) continue;
        PatternInfo e = v[i];
        e.str[a] = '?';
        v.push_back(e);
        if (b != a) {
            e.str[b] = '?';
            v.push_back(e);
            e.str[a] = rival;
            v.push_back(e);
        }
    }
    return v;
}

// ---- stage 2: ordering (ACAutomata.cpp:66-90) ----------------------------------------------
// Same key arithmetic, and std::sort itself over the same element type with the same
// comparison, so ties fall wherever they fall for the reference built with this toolchain.
void order_like_reference(std::vector<PatternInfo>& v) {
    std::vector<int> key(v.size()), idx(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
        int acc = 0;
        for (char ch : v[i].str) acc = acc * 4 + code_of(ch);
        key[i] = static_cast<int>(acc * std::pow(4.0, 7 - static_cast<int>(v[i].str.size())));
    }
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&key](int l, int r) { return key[l] < key[r]; });
    std::vector<PatternInfo> sorted;
    sorted.reserve(v.size());
    for (int i : idx) sorted.push_back(v[i]);
    v.swap(sorted);
}

// ---- stage 3: the interval-keyed trie (ACAutomata.cpp:105-134, ACAutomata.h:11-29,61-65) ----
// Nodes are identified by (depth, index of the first pattern below them); a node's children are
// whatever depth+1 nodes fall into its pattern interval.  Key collisions silently alias nodes.
struct IntervalTrie {
    struct Node { int code; int last; };
    using Key = std::pair<int, int>;                          // (depth, first)
    std::map<Key, Node> nodes;
    using It = std::map<Key, Node>::iterator;

    std::pair<It, It> children(It n) {
        return { nodes.lower_bound({ n->first.first + 1, n->first.second }),
                 nodes.upper_bound({ n->first.first + 1, n->second.last - 1 }) };
    }
    void add(It parent, const std::string& s, size_t at) {
        const int depth = parent->first.first + 1;
        if (at == s.size()) {                                 // leaf sentinel, code 0
            const int first = parent->first.second;
            const int last = ++parent->second.last;
            nodes.emplace(Key{ depth, first }, Node{ 0, last });
            return;
        }
        const int code = code_of(s[at]);
        auto [lo, hi] = children(parent);
        It child = std::find_if(lo, hi, [code](const auto& kv) { return kv.second.code == code; });
        if (child == hi)
            child = nodes.emplace(Key{ depth, parent->second.last }, Node{ code, parent->second.last }).first;
        add(child, s, at + 1);
        parent->second.last = child->second.last;
    }
};

// ---- stage 4: slot placement + fail links (ACAutomata.cpp:158-274) --------------------------
// Slots matter only because placement decides which (parent, code) pairs resolve to which
// node when the trie has aliased nodes; the result is read back as goto/fail/terminal below.
struct SlotAutomaton {
    std::vector<int> base, check, fail;
    int inv[5] = { 0, 0, 0, 0, 0 };
    bool ok = true;

    void grow_to(int need) {
        while (need >= static_cast<int>(check.size())) {
            const int old = static_cast<int>(base.size());
            base.resize(2 * old);
            check.resize(2 * old);
            for (int i = old; i < 2 * old; ++i) { base[i] = -(i - 1); check[i] = -(i + 1); }
        }
    }
    void place(IntervalTrie& trie, int slot, IntervalTrie::It node) {
        if (node->first.first > 0 && node->second.code == 0) { base[slot] = -node->first.second; return; }
        auto [lo, hi] = trie.children(node);
        if (lo == hi) { ok = false; return; }
        int begin = 0, front = 0;
        bool fits;
        do {
            front = -check[front];
            begin = front - lo->second.code;
            if (begin >= 0) grow_to(begin + 5);
            fits = true;
            for (auto c = lo; c != hi; ++c) {
                const int s = begin + c->second.code;
                if (s < 0 || s >= static_cast<int>(check.size())) { ok = false; return; }
                if (s == 0 || check[s] >= 0) { fits = false; break; }
            }
        } while (!fits);
        for (auto c = lo; c != hi; ++c) {
            const int s = begin + c->second.code;
            check[-base[s]] = check[s];
            base[-check[s]] = base[s];
            check[s] = slot;
        }
        base[slot] = begin;
        for (auto c = lo; c != hi; ++c) place(trie, begin + c->second.code, c);
    }
    void link() {
        fail.assign(base.size(), 0);
        std::deque<int> todo{ 0 };
        while (!todo.empty()) {
            const int cur = todo.front();
            todo.pop_front();
            for (int code = 1; code <= 4; ++code)
                if (check[base[cur] + code] == cur) todo.push_back(base[cur] + code);
            if (cur == 0) continue;
            const int code = cur - base[check[cur]];
            for (int up = check[cur]; up != 0;) {
                up = fail[up];
                const int cand = base[up] + code;
                if (check[cand] == up) { fail[cur] = cand; break; }
            }
            if (check[base[cur] + code] != cur && base[fail[cur]] + code == cur) inv[code] = cur;
        }
    }
    bool has(int s, int code) const { return check[base[s] + code] == s; }
    bool terminal(int s) const { return check[base[s]] == s; }
    int pattern(int s) const { return -base[base[s]]; }
};

// ---- stage 5: flatten ------------------------------------------------------------------------
struct Emit { int pid; bool prev; };
struct FlatStep { int next; std::vector<Emit> emits; };   // next: >= 0 slot, or -(slot) - 1 = "in a run at slot"

// One symbol of PatternSearch::generator::operator++ (Pattern.cpp:33-56) from slot s.
// in_run: s is a terminal invariant state and at least one extra symbol of its run has been
// skipped, i.e. one emission of pattern(s) is pending until the run ends.
FlatStep simulate(const SlotAutomaton& a, int s, bool in_run, int code) {
    FlatStep r{ 0, {} };
    if (in_run) {
        if (a.inv[code] == s) { r.next = -s - 1; return r; }
        r.emits.push_back({ a.pattern(s), true });            // the run ended on the previous symbol
    }
    for (;;) {
        if (s == a.inv[code]) {                               // run skip: symbol consumed, state kept
            r.next = a.terminal(s) ? -s - 1 : s;
            return r;
        }
        if (a.has(s, code)) {
            s = a.base[s] + code;
            if (a.terminal(s)) r.emits.push_back({ a.pattern(s), false });
            r.next = s;
            return r;
        }
        if (s == 0) { r.next = 0; return r; }                 // mismatch at the root: skip the symbol
        s = a.fail[s];
        if (a.terminal(s)) r.emits.push_back({ a.pattern(s), true });
    }
}

// device pattern record (gk_format.h); false if the pattern scores more than four cells
bool make_patrec(const PatternInfo& p, PatRec& out) {
    uint32_t w0 = 0;
    int n = 0;
    const int len = static_cast<int>(p.str.size());
    for (int j = 0; j < len; ++j) {
        const char ch = p.str[len - 1 - j];
        if (ch != '_' && ch != '^') continue;
        if (n == 4) return false;
        w0 |= static_cast<uint32_t>(j | (ch == '_' ? 8 : 0)) << (4 * n++);
    }
    w0 |= static_cast<uint32_t>(n) << 16;
    w0 |= static_cast<uint32_t>(p.type) << 19;
    w0 |= static_cast<uint32_t>(p.favour == 1) << 23;
    w0 |= static_cast<uint32_t>(len) << 24;
    const int cclass = p.type == 5 ? 1 : p.type == 4 ? 2 : p.type == 3 ? 3 : 0;   // CompTypes, Pattern.cpp:420-422
    w0 |= static_cast<uint32_t>(cclass) << 27;
    const int diag = static_cast<int>(1 * 1.2 * p.score);     // Pattern.cpp:151-152 with delta = +1
    out = { w0, static_cast<uint32_t>(p.score) | static_cast<uint32_t>(diag) << 16 };
    return true;
}

struct Line { int cell0, stride, len, dir; };
std::vector<Line> board_lines() {                              // the 72 lines that can hold a pattern
    std::vector<Line> v;
    for (int y = 0; y < kHeight; ++y) v.push_back({ y * kWidth, 1, kWidth, 0 });
    for (int x = 0; x < kWidth; ++x) v.push_back({ x, kWidth, kHeight, 1 });
    for (int k = -(kHeight - 1); k < kWidth; ++k) {            // x - y = k, walking (+1,+1)
        const int len = kWidth - std::abs(k);
        if (len >= 5) v.push_back({ (k > 0 ? 0 : -k) * kWidth + (k > 0 ? k : 0), kWidth + 1, len, 2 });
    }
    for (int k = 0; k < kWidth + kHeight - 1; ++k) {           // x + y = k, walking (-1,+1)
        const int x0 = std::min(k, kWidth - 1), len = (k < kWidth ? k : 2 * (kWidth - 1) - k) + 1;
        if (len >= 5) v.push_back({ (k - x0) * kWidth + x0, kWidth - 1, len, 3 });
    }
    return v;
}

// Longest-processing-time-first packing of whole lines onto the 32 lanes of a warp.
bool build_tape(HostTable& t) {
    std::vector<Line> lines = board_lines();
    std::vector<int> order(lines.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lines[a].len > lines[b].len; });
    std::vector<std::vector<int>> lane(32);
    std::vector<int> load(32, 0);
    for (int li : order) {
        const int best = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
        lane[best].push_back(li);
        load[best] += lines[li].len + t.trail_pad;
    }
    t.tape_steps = (*std::max_element(load.begin(), load.end()) + 1) & ~1;   // the kernel unrolls by two
    if (t.tape_steps > kMaxTapeSteps) return false;
    const auto src_of = [](int cell) {                        // gk_format.h: board word index | rotate amount << 8
        return static_cast<uint16_t>((cell >> 4) | (((2 * (cell & 15) + 31) & 31) << 8));
    };
    t.tape_src.assign(static_cast<size_t>(t.tape_steps) * 32, src_of(kPadCell));   // filler: a pad that belongs to no line
    t.tape_info.assign(static_cast<size_t>(t.tape_steps) * 32, 0);
    for (int l = 0; l < 32; ++l) {
        int step = 0;
        for (int li : lane[l]) {
            const Line& L = lines[li];
            for (int i = 0; i < L.len + t.trail_pad; ++i, ++step) {
                const int vcell = L.cell0 + i * L.stride;
                t.tape_src[static_cast<size_t>(step) * 32 + l] = src_of(i < L.len ? vcell : kPadCell);
                t.tape_info[static_cast<size_t>(step) * 32 + l] = static_cast<uint16_t>(vcell | L.dir << 9 | L.stride << 11);
            }
        }
    }
    // Emission-list capacity: the largest number of emitting steps any lane chain can produce, over ALL
    // boards (cells take the values empty / black / white, pads are fixed) -- a max-plus walk of the
    // automaton along each lane's tape.  The kernel's per-lane lists are sized by this, so they cannot overflow.
    int worst = 0;
    for (int l = 0; l < 32; ++l) {
        std::vector<int> best(t.n_states, -1), next_best(t.n_states);
        best[t.start_state] = 0;
        for (int step = 0; step < t.tape_steps; ++step) {
            const bool pad = t.tape_src[static_cast<size_t>(step) * 32 + l] == src_of(kPadCell);
            std::fill(next_best.begin(), next_best.end(), -1);
            for (int st = 0; st < t.n_states; ++st) {
                if (best[st] < 0) continue;
                for (int sym = 0; sym < 4; ++sym) {
                    if (pad != (sym == kSymPad)) continue;
                    const uint32_t w = t.trans[static_cast<size_t>(st) * 4 + sym];
                    const int got = best[st] + (tw_nemit(w) != 0 ? 1 : 0);
                    int& slot = next_best[tw_next(w)];
                    slot = std::max(slot, got);
                }
            }
            best.swap(next_best);
        }
        worst = std::max(worst, *std::max_element(best.begin(), best.end()));
    }
    t.list_cap = std::max(worst, 16);
    while (t.list_cap % 4 != 2) ++t.list_cap;                  // cap / 2 odd: lanes' lists start in different banks
    return true;
}

}  // namespace

bool compile_table(const std::vector<Proto>& protos, HostTable& t) {
    t = HostTable{};
    for (const Proto& p : protos) {
        if (p.text.size() < 2 || p.text.size() > 8 || (p.text[0] != '+' && p.text[0] != '-') || p.type < 0 ||
            p.type > kTypeFive || p.score < 0 || p.score * 1.2 > 65535) {
            t.error = "bad prototype '" + p.text + "'";
            return false;
        }
        for (size_t i = 1; i < p.text.size(); ++i)
            if (code_of(p.text[i]) == 0) { t.error = "bad symbol in prototype '" + p.text + "'"; return false; }
    }
    t.patterns = expand(protos);
    if (t.patterns.empty() || static_cast<int>(t.patterns.size()) > kMaxPatterns) { t.error = "pattern count out of range"; return false; }
    order_like_reference(t.patterns);

    IntervalTrie trie;
    auto root = trie.nodes.emplace(IntervalTrie::Key{ 0, 0 }, IntervalTrie::Node{ 0, 0 }).first;
    for (const PatternInfo& p : t.patterns) trie.add(root, p.str, 0);

    SlotAutomaton a;
    a.base.assign(1, 0);
    a.check.assign(1, -1);
    a.place(trie, 0, root);
    if (!a.ok) { t.error = "prototype set breaks the reference's trie construction"; return false; }
    a.link();

    // dense ids: breadth-first over goto edges, then one "in run" twin per terminal invariant
    std::map<int, int> dense;                                  // key: slot, or -(slot)-1 for a twin
    std::vector<int> slots;
    {
        std::deque<int> todo{ 0 };
        dense[0] = 0;
        slots.push_back(0);
        while (!todo.empty()) {
            const int s = todo.front();
            todo.pop_front();
            for (int code = 1; code <= 4; ++code)
                if (a.has(s, code)) {
                    const int c = a.base[s] + code;
                    if (dense.emplace(c, static_cast<int>(slots.size())).second) { slots.push_back(c); todo.push_back(c); }
                }
        }
        for (int code = 1; code <= 4; ++code) {
            const int s = a.inv[code];
            if (s != 0 && dense.count(s) && a.terminal(s) && dense.emplace(-s - 1, static_cast<int>(slots.size())).second)
                slots.push_back(-s - 1);
        }
    }
    t.n_states = static_cast<int>(slots.size());
    if (t.n_states > kMaxStates) { t.error = "too many automaton states"; return false; }
    t.trans.assign(static_cast<size_t>(t.n_states) * 4, 0);
    t.flush.assign(t.n_states, -1);
    for (int id = 0; id < t.n_states; ++id) {
        const bool twin = slots[id] < 0;
        const int s = twin ? -slots[id] - 1 : slots[id];
        if (twin) t.flush[id] = static_cast<int16_t>(a.pattern(s));
        for (int code = 1; code <= 4; ++code) {
            const FlatStep st = simulate(a, s, twin, code);
            if (st.emits.size() > 2) { t.error = "more than two emissions on one transition"; return false; }
            auto it = dense.find(st.next);
            if (it == dense.end()) { t.error = "transition leaves the reachable state set"; return false; }
            uint32_t w = static_cast<uint32_t>(it->second) | static_cast<uint32_t>(st.emits.size()) << 10;
            for (size_t k = 0; k < st.emits.size(); ++k)
                w |= (static_cast<uint32_t>(st.emits[k].pid) | static_cast<uint32_t>(st.emits[k].prev) << 9) << (12 + 10 * k);
            t.trans[static_cast<size_t>(id) * 4 + sym_from_refcode(code)] = w;
        }
    }
    for (const PatternInfo& p : t.patterns) {
        PatRec rec;
        if (!make_patrec(p, rec)) { t.error = "pattern '" + p.str + "' scores more than four cells"; return false; }
        t.patrec.push_back(rec);
    }
    if (static_cast<int>(t.patterns.size()) >= static_cast<int>(kDevNoPid)) { t.error = "pattern id 511 is reserved"; return false; }

    // A board line is "?", cells, "?..." : after the leading pad the automaton must sit in a
    // state that further pads do not move, so the kernel can start every line there ...
    {
        const uint32_t w = t.trans[kSymPad];
        t.start_state = static_cast<int>(tw_next(w));
        const uint32_t again = t.trans[static_cast<size_t>(t.start_state) * 4 + kSymPad];
        if (tw_nemit(w) != 0 || tw_nemit(again) != 0 || static_cast<int>(tw_next(again)) != t.start_state) {
            t.error = "leading board edge is not a fixed point of the automaton";
            return false;
        }
    }
    // ... and behind the last cell, pads are fed until no state can emit any more.
    {
        std::vector<int> cur(t.n_states);
        std::iota(cur.begin(), cur.end(), 0);
        int last_emitting = 0;
        for (int k = 1; k <= 16; ++k) {
            bool any = false;
            for (int& s : cur) {
                const uint32_t w = t.trans[static_cast<size_t>(s) * 4 + kSymPad];
                any = any || tw_nemit(w) != 0;
                s = static_cast<int>(tw_next(w));
            }
            if (any) last_emitting = k;
        }
        if (last_emitting >= 16) { t.error = "trailing board edge never stops emitting"; return false; }
        t.trail_pad = std::max(last_emitting, 2);
        // trail_pad pads must also leave every state in start_state: that is what lets the kernel run the
        // lines of a lane back to back without restarting the automaton.
        std::iota(cur.begin(), cur.end(), 0);
        for (int k = 0; k < t.trail_pad; ++k)
            for (int& s : cur) s = static_cast<int>(tw_next(t.trans[static_cast<size_t>(s) * 4 + kSymPad]));
        for (int s : cur)
            if (s != t.start_state) { t.error = "board edge does not reset the automaton"; return false; }
    }
    // device automaton (gk_format.h): one clone of the destination per distinct (destination, emission set)
    {
        std::map<std::pair<int, uint32_t>, int> clone_of;     // (destination, emission bits of the host word) -> clone id
        std::vector<std::pair<int, uint32_t>> clones;
        for (int id = 0; id < t.n_states; ++id)
            for (int sym = 0; sym < 4; ++sym) {
                const uint32_t w = t.trans[static_cast<size_t>(id) * 4 + sym];
                if (tw_nemit(w) == 0) continue;
                const auto key = std::make_pair(static_cast<int>(tw_next(w)), w >> 10);
                if (clone_of.emplace(key, static_cast<int>(clones.size())).second) clones.push_back(key);
            }
        t.n_clones = static_cast<int>(clones.size());
        const int total = t.n_clones + t.n_states;
        if (t.n_clones > kMaxClones || total * 8 > 65535) { t.error = "automaton too large for the 16-bit device table"; return false; }
        t.dev_next.assign(static_cast<size_t>(total) * 4, 0);
        t.dev_erec.assign(t.n_clones, 0);
        const auto row_of = [&](int dev_id) { return dev_id < t.n_clones ? clones[dev_id].first : dev_id - t.n_clones; };
        for (int dev_id = 0; dev_id < total; ++dev_id)
            for (int sym = 0; sym < 4; ++sym) {
                const uint32_t w = t.trans[static_cast<size_t>(row_of(dev_id)) * 4 + sym];
                const int dest = tw_nemit(w) == 0 ? t.n_clones + static_cast<int>(tw_next(w))
                                                  : clone_of.at({ static_cast<int>(tw_next(w)), w >> 10 });
                t.dev_next[static_cast<size_t>(dev_id) * 4 + sym_to_value(sym)] = static_cast<uint16_t>(dest * 8);
            }
        for (int c = 0; c < t.n_clones; ++c) {
            const uint32_t bits = clones[c].second;            // host word >> 10: count, emission 0, emission 1
            const int n = static_cast<int>(bits & 3u);
            uint32_t e = 0;
            for (int k = 0; k < 2; ++k) {
                const uint32_t em = (bits >> (2 + 10 * k)) & 0x3ffu;
                e |= (k < n ? (em_pid(em) | em_prev(em) << 9) : kDevNoPid) << (10 * k);
            }
            const int type = t.patterns[em_pid((bits >> 2) & 0x3ffu)].type;
            e |= static_cast<uint32_t>(type == 5 ? 1 : type == 4 ? 2 : type == 3 ? 3 : 0) << 20;
            t.dev_erec[c] = e;
        }
        t.root_off = t.n_clones * 8;
        t.start_off = (t.n_clones + t.start_state) * 8;
    }
    // How many symbols until the state forgets where it started (0: it never does).  Informational:
    // it bounds the context an emission can depend on.
    {
        std::vector<std::vector<int>> image{ std::vector<int>(t.n_states) };
        std::iota(image[0].begin(), image[0].end(), 0);
        for (int depth = 1; depth <= 8 && t.sync_depth == 0; ++depth) {
            std::vector<std::vector<int>> next;
            bool all_single = true;
            for (const auto& set : image)
                for (int sym = 0; sym < 4; ++sym) {
                    std::vector<int> img;
                    for (int s : set) img.push_back(static_cast<int>(tw_next(t.trans[static_cast<size_t>(s) * 4 + sym])));
                    std::sort(img.begin(), img.end());
                    img.erase(std::unique(img.begin(), img.end()), img.end());
                    all_single = all_single && img.size() == 1;
                    next.push_back(std::move(img));
                }
            image.swap(next);
            if (all_single) t.sync_depth = depth;
        }
    }
    if (!build_tape(t)) { t.error = "scan tape longer than 64 steps"; return false; }
    return true;
}

}  // namespace gk
