/* ORACLE -- test infrastructure, not product code.  See gomoku_oracle.h.
 *
 * Plain-C restatement of the reference's hot path.  All paths cited below are relative
 * to /root/reference/core/lib/.  The structure deliberately mirrors the reference
 * (double-array trie, generator-style matcher, incremental evaluator with a -1 phase and
 * a +1 phase per move) so that it can be compared array-by-array with the reference
 * compiled in oracle/_ref/.  It shares no code with gomokuai_b200/csrc, which uses a flat
 * transducer table and a from-scratch, per-position evaluation.
 */
#include "gomoku_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ======================================================================================
 * Charset and players            include/Mapping.h:40-51, include/Game.h:19-36
 * ==================================================================================== */
enum { P_WHITE = -1, P_NONE = 0, P_BLACK = 1 };
enum { T_DEAD1, T_LIVE1, T_DEAD2, T_LIVE2, T_DEAD3, T_LIVE3, T_DEAD4, T_LIVE4, T_FIVE, T_SIZE };
enum { C_33, C_43, C_44, C_SIZE };
enum { D_H, D_V, D_LD, D_RD };

/* EncodeCharset, Mapping.h:40-48: x=1 o=2 ?=3 blank(- _ ^ ~)=4, anything else 0 */
int orc_encode(char ch) {
    switch (ch) {
        case '-': case '_': case '^': case '~': return 4;
        case '?': return 3;
        case 'o': return 2;
        case 'x': return 1;
        default: return 0;
    }
}

/* operator*(Direction), Mapping.h:14-22 */
static const int DIR_DX[4] = { 1, 0, 1, -1 };
static const int DIR_DY[4] = { 0, 1, 1, 1 };
/* Shift(), Mapping.h:25-27: id + offset * Position(dx, dy).id */
static int shift(int id, int offset, int dir) { return id + offset * (DIR_DY[dir] * ORC_W + DIR_DX[dir]); }

/* ======================================================================================
 * Pattern table                  src/Pattern.cpp:14-18 (ctor), :554-596 (data)
 * ==================================================================================== */
typedef struct { char str[8]; int len; int favour; int type; int score; } pat_t;

static void pat_from_proto(pat_t* p, const char* proto, int type, int score) {
    memset(p, 0, sizeof *p);
    p->len = (int)strlen(proto) - 1;
    memcpy(p->str, proto + 1, (size_t)p->len);
    p->favour = proto[0] == '+' ? P_BLACK : P_WHITE;
    p->type = type;
    p->score = score;
}

static const struct { const char* proto; int type; int score; } DEFAULT_PROTOS[] = {
    { "+xxxxx",   T_FIVE,  9999 }, { "-_oooo_",  T_LIVE4, 9000 }, { "-xoooo_",  T_DEAD4, 2500 },
    { "-o_ooo",   T_DEAD4, 3000 }, { "-oo_oo",   T_DEAD4, 2600 }, { "-~_ooo_~", T_LIVE3, 3000 },
    { "-x^ooo_~", T_LIVE3, 2900 }, { "-~o_oo~",  T_LIVE3, 2800 }, { "-~o~oo_~", T_DEAD3, 1400 },
    { "-~oo~o_~", T_DEAD3, 1200 }, { "-x_o~oo~", T_DEAD3, 1300 }, { "-x_oo~o~", T_DEAD3, 1100 },
    { "-xooo__~", T_DEAD3, 510 },  { "-xoo_o_~", T_DEAD3, 520 },  { "-xoo__o~", T_DEAD3, 520 },
    { "-xo_oo_~", T_DEAD3, 530 },  { "-xo__oo",  T_DEAD3, 530 },  { "-xooo__x", T_DEAD3, 500 },
    { "-xoo_o_x", T_DEAD3, 500 },  { "-xoo__ox", T_DEAD3, 500 },  { "-xo_oo_x", T_DEAD3, 500 },
    { "-x_ooo_x", T_DEAD3, 500 },  { "-~oo__o~", T_DEAD3, 750 },  { "-oo__oo",  T_DEAD3, 540 },
    { "-o_o_o",   T_DEAD3, 550 },  { "-~oo__~",  T_LIVE2, 650 },  { "-~_o_o_~", T_LIVE2, 600 },
    { "-x^o_o_^", T_LIVE2, 550 },  { "-^o__o^",  T_LIVE2, 550 },  { "-xoo___",  T_DEAD2, 150 },
    { "-xo_o__",  T_DEAD2, 160 },  { "-xo__o_",  T_DEAD2, 170 },  { "-o___o",   T_DEAD2, 180 },
    { "-x_oo__x", T_DEAD2, 120 },  { "-x_o_o_x", T_DEAD2, 120 },  { "-~o___~",  T_LIVE1, 150 },
    { "-x~_o__^", T_LIVE1, 140 },  { "-x~__o_^", T_LIVE1, 150 },  { "-xo___~",  T_DEAD1, 30 },
    { "-x_o___x", T_DEAD1, 40 },   { "-x__o__x", T_DEAD1, 50 },
};
enum { N_DEFAULT_PROTOS = sizeof DEFAULT_PROTOS / sizeof DEFAULT_PROTOS[0] };

/* BlockWeights, src/Pattern.cpp:598-609 */
static const int BLOCK_W[7][7] = {
    { 2, 0, 0, 1, 0, 0, 2 }, { 0, 4, 3, 3, 3, 4, 0 }, { 0, 3, 5, 4, 5, 3, 0 }, { 1, 3, 4, 0, 4, 3, 1 },
    { 0, 3, 5, 4, 5, 3, 0 }, { 0, 4, 3, 3, 3, 4, 0 }, { 2, 0, 0, 1, 0, 0, 2 },
};
enum { BLOCK_SCORE = 160, COMPOUND_SCORE = 600 };   /* Pattern.cpp:600, :611 */

/* ======================================================================================
 * AhoCorasickBuilder             src/utils/ACAutomata.cpp
 * ==================================================================================== */
struct orc_table {
    pat_t pats[ORC_MAX_PATTERNS];
    int n_pats;
    int* base; int* check; int* fail; int size;
    int inv[5];
    int build_error;            /* non-zero if an assumption of the reference's builder broke */
};

/* reverseAugment, ACAutomata.cpp:25-33 */
static void aug_reverse(pat_t* v, int* n) {
    int size = *n;
    for (int i = 0; i < size; ++i) {
        pat_t r = v[i];
        for (int k = 0; k < r.len; ++k) r.str[k] = v[i].str[r.len - 1 - k];
        if (memcmp(r.str, v[i].str, 8) != 0) v[(*n)++] = r;
    }
}
/* flipAugment, ACAutomata.cpp:35-45 */
static void aug_flip(pat_t* v, int* n) {
    int size = *n;
    for (int i = 0; i < size; ++i) {
        pat_t f = v[i];
        f.favour = -f.favour;
        for (int k = 0; k < f.len; ++k) {
            if (f.str[k] == 'x') f.str[k] = 'o';
            else if (f.str[k] == 'o') f.str[k] = 'x';
        }
        v[(*n)++] = f;
    }
}
/* boundaryAugment, ACAutomata.cpp:47-64 */
static void aug_boundary(pat_t* v, int* n) {
    int size = *n;
    for (int i = 0; i < size; ++i) {
        char enemy = v[i].favour == P_BLACK ? 'o' : 'x';
        int first = -1, last = -1;
        for (int k = 0; k < v[i].len; ++k)
            if (v[i].str[k] == enemy) { if (first < 0) first = k; last = k; }
        if (first >= 0) {
            pat_t b = v[i];
            b.str[first] = '?';
            v[(*n)++] = b;
            if (last != first) {
                b.str[last] = '?';
                v[(*n)++] = b;
                b.str[first] = enemy;
                v[(*n)++] = b;
            }
        }
    }
}
/* sortPatterns, ACAutomata.cpp:66-90: key = base-4 accumulation of the codes (int), times
 * pow(4, 7 - len) (double), truncated to int; ascending by std::sort. */
static int sort_key(const pat_t* p) {
    double align = pow(4.0, (double)(ORC_MAX_PATTERN_LEN - p->len));
    int sum = 0;
    for (int k = 0; k < p->len; ++k) { sum *= 4; sum += orc_encode(p->str[k]); }
    return (int)(sum * align);
}
/* std::sort as implemented by libstdc++ (bits/stl_algo.h: __introsort_loop with a
 * median-of-3 pivot, threshold 16, then __final_insertion_sort).  The default table has
 * three pairs of equal keys ("-x--xx-" vs "-x--xo" etc.: bijective base-4 digits 1..4 make
 * 16*v+8 reachable two ways), so the order of those pairs -- and with it pattern ids and
 * the DAT slot numbering -- is whatever the standard library's unstable sort yields.  The
 * oracle is pinned against the reference compiled with g++/libstdc++ here, hence this. */
static const int* g_sort_keys;
static int key_less(int a, int b) { return g_sort_keys[a] < g_sort_keys[b]; }
static void iswap(int* a, int* b) { int t = *a; *a = *b; *b = t; }
static void sort_linear_insert(int* last) {
    int val = *last; int* next = last - 1;
    while (key_less(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void sort_insertion(int* first, int* last) {
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        if (key_less(*i, *first)) { int val = *i; memmove(first + 1, first, sizeof(int) * (size_t)(i - first)); *first = val; }
        else sort_linear_insert(i);
    }
}
static int sort_introloop(int* first, int* last, int depth) {
    while (last - first > 16) {
        if (depth == 0) return 1;                    /* heapsort fallback: not restated, flagged */
        --depth;
        int* mid = first + (last - first) / 2;
        int *a = first + 1, *b = mid, *c = last - 1; /* __move_median_to_first(first, a, b, c) */
        if (key_less(*a, *b)) {
            if (key_less(*b, *c)) iswap(first, b);
            else if (key_less(*a, *c)) iswap(first, c);
            else iswap(first, a);
        } else if (key_less(*a, *c)) iswap(first, a);
        else if (key_less(*b, *c)) iswap(first, c);
        else iswap(first, b);
        int *lo = first + 1, *hi = last;             /* __unguarded_partition(first + 1, last, first) */
        for (;;) {
            while (key_less(*lo, *first)) ++lo;
            --hi;
            while (key_less(*first, *hi)) --hi;
            if (!(lo < hi)) break;
            iswap(lo, hi);
            ++lo;
        }
        if (sort_introloop(lo, last, depth)) return 1;
        last = lo;
    }
    return 0;
}
static void aug_sort(pat_t* v, int n, int* err) {
    int* key = (int*)malloc(sizeof(int) * (size_t)n);
    int* idx = (int*)malloc(sizeof(int) * (size_t)n);
    pat_t* tmp = (pat_t*)malloc(sizeof(pat_t) * (size_t)n);
    for (int i = 0; i < n; ++i) { key[i] = sort_key(&v[i]); idx[i] = i; }
    g_sort_keys = key;
    if (n > 0) {
        int lg = 0; while ((1 << (lg + 1)) <= n) ++lg;
        if (sort_introloop(idx, idx + n, 2 * lg) && err) *err = 1;
        if (n > 16) { sort_insertion(idx, idx + 16); for (int* i = idx + 16; i != idx + n; ++i) sort_linear_insert(i); }
        else sort_insertion(idx, idx + n);
    }
    for (int i = 0; i < n; ++i) tmp[i] = v[idx[i]];
    memcpy(v, tmp, sizeof(pat_t) * (size_t)n);
    free(key); free(idx); free(tmp);
}

/* buildNodeBasedTrie, ACAutomata.cpp:105-134.  The reference keeps trie nodes in a
 * std::set<Node> ordered by (depth, first) ONLY (ACAutomata.h:22-24); `last` is mutable and
 * the children of a node are found by a RANGE QUERY: the depth+1 nodes whose `first` lies in
 * [node.first, node.last - 1] (ACAutomata.h:61-65).  This must be restated literally, not as
 * a pointer trie: where sortPatterns leaves two equal keys in the "wrong" order (see above)
 * a new branch node can collide with an existing (depth, first) key, std::set::insert then
 * hands back the EXISTING node (a leaf sentinel of the other pattern), and the resulting
 * automaton reports pattern 217 where 218 ("-x--xx-") matches, etc.  The reference built with
 * libstdc++ has exactly that automaton, so the oracle reproduces it. */
typedef struct { int code, depth, first, last; } tnode;
typedef struct { tnode pool[8192]; int order[8192]; int n; } tset;    /* order[]: pool indices sorted by (depth, first) */

static int tset_less(const tnode* a, int depth, int first) { return a->depth < depth || (a->depth == depth && a->first < first); }
static int tset_lower(const tset* s, int depth, int first) {          /* first position with key >= (depth, first) */
    int lo = 0, hi = s->n;
    while (lo < hi) { int mid = (lo + hi) / 2; if (tset_less(&s->pool[s->order[mid]], depth, first)) lo = mid + 1; else hi = mid; }
    return lo;
}
static int tset_upper(const tset* s, int depth, int first) { return tset_lower(s, depth, first + 1); }   /* integer keys */
/* std::set::insert / emplace: returns the pool index of the new node, or of the existing
 * node with an equal key (in which case nothing is inserted). */
static int tset_insert(tset* s, tnode nd) {
    int pos = tset_lower(s, nd.depth, nd.first);
    if (pos < s->n) {
        const tnode* e = &s->pool[s->order[pos]];
        if (e->depth == nd.depth && e->first == nd.first) return s->order[pos];
    }
    if (s->n >= 8192) return -1;
    int id = s->n;
    s->pool[id] = nd;
    memmove(&s->order[pos + 1], &s->order[pos], sizeof(int) * (size_t)(s->n - pos));
    s->order[pos] = id;
    s->n += 1;
    return id;
}
/* children(), ACAutomata.h:61-65 -> [lo, hi) positions in order[] */
static void tset_children(const tset* s, const tnode* node, int* lo, int* hi) {
    *lo = tset_lower(s, node->depth + 1, node->first);
    *hi = tset_upper(s, node->depth + 1, node->last - 1);
}
static void trie_insert(tset* s, int parent, const char* suffix, int len, int* err) {
    if (len == 0) {                       /* ACAutomata.cpp:108-111: sentinel leaf, code 0 */
        tnode leaf; leaf.code = 0; leaf.depth = s->pool[parent].depth + 1;
        leaf.first = s->pool[parent].first; leaf.last = ++s->pool[parent].last;
        if (tset_insert(s, leaf) < 0) *err = 1;
        return;
    }
    int code = orc_encode(suffix[0]), lo, hi, child = -1;
    tset_children(s, &s->pool[parent], &lo, &hi);
    for (int i = lo; i < hi; ++i)
        if (s->pool[s->order[i]].code == code) { child = s->order[i]; break; }
    if (child < 0) {                      /* :121-125 new branch with an empty range at parent->last */
        tnode key; key.code = code; key.depth = s->pool[parent].depth + 1;
        key.first = s->pool[parent].last; key.last = key.first;
        child = tset_insert(s, key);      /* may return an EXISTING node, see header comment */
        if (child < 0) { *err = 1; return; }
    }
    trie_insert(s, child, suffix + 1, len - 1, err);
    s->pool[parent].last = s->pool[child].last;      /* :127 */
}

/* buildDAT, ACAutomata.cpp:158-229 (free slots form a doubly linked list threaded through
 * negative base/check values; slot 0's check is the list head). */
static void dat_grow(orc_table* t, int need_index) {
    /* ACAutomata.cpp:185-195: double until begin + |Codeset| + 1 < size */
    while (need_index >= t->size) {
        int pre = t->size, now = 2 * pre;
        t->base = (int*)realloc(t->base, sizeof(int) * (size_t)now);
        t->check = (int*)realloc(t->check, sizeof(int) * (size_t)now);
        for (int i = pre; i < now; ++i) { t->base[i] = -(i - 1); t->check[i] = -(i + 1); }
        t->size = now;
    }
}
static void dat_build(orc_table* t, const tset* s, int index, const tnode* node) {
    if (node->depth > 0 && node->code == 0) {        /* :161-164 leaf -> -(pattern index) */
        t->base[index] = -node->first;
        return;
    }
    int lo, hi;
    tset_children(s, node, &lo, &hi);
    if (lo >= hi) { t->build_error = 4; return; }    /* the reference would dereference end() */
    int begin = 0, front = 0, ok;
    do {                                             /* :176-202 */
        front = -t->check[front];
        begin = front - s->pool[s->order[lo]].code;
        if (begin >= 0) dat_grow(t, begin + 4 + 1);   /* `continue` at :181 skips only the growth */
        ok = 1;
        for (int i = lo; i < hi; ++i) {
            int ci = begin + s->pool[s->order[i]].code;
            if (ci < 0 || ci >= t->size) { t->build_error = 2; return; }   /* UB in the reference */
            if (!(ci != 0 && t->check[ci] < 0)) { ok = 0; break; }
        }
    } while (!ok);
    for (int i = lo; i < hi; ++i) {                  /* :205-215 unlink from the free list, bind */
        int ci = begin + s->pool[s->order[i]].code;
        t->check[-t->base[ci]] = t->check[ci];
        t->base[-t->check[ci]] = t->base[ci];
        t->check[ci] = index;
    }
    t->base[index] = begin;                          /* :217 */
    for (int i = lo; i < hi; ++i) {
        const tnode* c = &s->pool[s->order[i]];
        dat_build(t, s, begin + c->code, c);
    }
}

/* buildACGraph, ACAutomata.cpp:231-274 */
static void ac_build(orc_table* t) {
    t->fail = (int*)calloc((size_t)t->size, sizeof(int));
    memset(t->inv, 0, sizeof t->inv);
    int* queue = (int*)malloc(sizeof(int) * (size_t)t->size);
    int qh = 0, qt = 0;
    queue[qt++] = 0;
    while (qh < qt) {
        int cur = queue[qh++];
        for (int code = 1; code <= 4; ++code) {
            int child = t->base[cur] + code;
            if (t->check[child] == cur) queue[qt++] = child;
        }
        if (cur == 0) continue;
        int code = cur - t->base[t->check[cur]];
        int pre = t->check[cur];
        while (pre != 0) {
            pre = t->fail[pre];
            int f = t->base[pre] + code;
            if (t->check[f] == pre) { t->fail[cur] = f; break; }
        }
        if (t->check[t->base[cur] + code] != cur && t->base[t->fail[cur]] + code == cur)
            t->inv[code] = cur;                      /* :269-272 "invariant" self-loop state */
    }
    free(queue);
}

/* AhoCorasickBuilder::build, ACAutomata.cpp:15-23 */
static orc_table* table_from_patterns(const pat_t* protos, int n) {
    orc_table* t = (orc_table*)calloc(1, sizeof *t);
    memcpy(t->pats, protos, sizeof(pat_t) * (size_t)n);
    t->n_pats = n;
    aug_reverse(t->pats, &t->n_pats);
    aug_flip(t->pats, &t->n_pats);
    aug_boundary(t->pats, &t->n_pats);
    int sort_err = 0;
    aug_sort(t->pats, t->n_pats, &sort_err);
    if (sort_err) t->build_error = 3;
    tset* set = (tset*)calloc(1, sizeof *set);
    tnode rootn = { 0, 0, 0, 0 };
    int root = tset_insert(set, rootn), err = 0;
    for (int i = 0; i < t->n_pats; ++i) trie_insert(set, root, t->pats[i].str, t->pats[i].len, &err);
    if (err) t->build_error = 1;
    t->size = 1;
    t->base = (int*)malloc(sizeof(int));
    t->check = (int*)malloc(sizeof(int));
    t->base[0] = 0; t->check[0] = -1;                /* :224-225 */
    dat_build(t, set, 0, &set->pool[root]);
    free(set);
    ac_build(t);
    return t;
}

orc_table* orc_table_build(const char* const* protos, const int* types, const int* scores, int n) {
    if (n * 12 > ORC_MAX_PATTERNS) return NULL;
    pat_t* v = (pat_t*)calloc((size_t)n, sizeof(pat_t));
    for (int i = 0; i < n; ++i) pat_from_proto(&v[i], protos[i], types[i], scores[i]);
    orc_table* t = table_from_patterns(v, n);
    free(v);
    return t;
}

orc_table* orc_table_default(void) {
    static orc_table* cached = NULL;                 /* Evaluator::Patterns, Pattern.cpp:554 */
    if (!cached) {
        pat_t v[N_DEFAULT_PROTOS];
        for (int i = 0; i < N_DEFAULT_PROTOS; ++i)
            pat_from_proto(&v[i], DEFAULT_PROTOS[i].proto, DEFAULT_PROTOS[i].type, DEFAULT_PROTOS[i].score);
        cached = table_from_patterns(v, N_DEFAULT_PROTOS);
    }
    return cached;
}

void orc_table_free(orc_table* t) {
    if (!t || t == orc_table_default()) return;
    free(t->base); free(t->check); free(t->fail); free(t);
}

int orc_table_sizes(const orc_table* t, int* n_base, int* n_patterns) {
    *n_base = t->size; *n_patterns = t->n_pats;
    return t->build_error;
}
int orc_table_arrays(const orc_table* t, int32_t* base, int32_t* check, int32_t* fail, int32_t* invariants) {
    for (int i = 0; i < t->size; ++i) { base[i] = t->base[i]; check[i] = t->check[i]; fail[i] = t->fail[i]; }
    for (int i = 0; i < 5; ++i) invariants[i] = t->inv[i];
    return 0;
}
int orc_table_pattern(const orc_table* t, int id, char* str8, int* favour, int* type, int* score) {
    if (id < 0 || id >= t->n_pats) return -1;
    memcpy(str8, t->pats[id].str, 8);
    *favour = t->pats[id].favour; *type = t->pats[id].type; *score = t->pats[id].score;
    return 0;
}
int orc_augment(const char* const* protos, const int* types, const int* scores, int n, int stage,
                char* strs, int* favours, int* otypes, int* oscores, int cap) {
    pat_t* v = (pat_t*)calloc((size_t)n * 12 + 1, sizeof(pat_t));
    int m = n;
    for (int i = 0; i < n; ++i) pat_from_proto(&v[i], protos[i], types[i], scores[i]);
    if (stage >= 1) aug_reverse(v, &m);
    if (stage >= 2) aug_flip(v, &m);
    if (stage >= 3) aug_boundary(v, &m);
    if (stage >= 4) aug_sort(v, m, NULL);
    for (int i = 0; i < m && i < cap; ++i) {
        memcpy(strs + 8 * i, v[i].str, 8);
        favours[i] = v[i].favour; otypes[i] = v[i].type; oscores[i] = v[i].score;
    }
    free(v);
    return m;
}

/* ======================================================================================
 * PatternSearch::generator       src/Pattern.cpp:33-62, include/Pattern.h:61-74
 * ==================================================================================== */
typedef struct { const uint8_t* t; int n; int offset; int state; const orc_table* tb; } gen_t;

static gen_t gen_make(const orc_table* tb, const uint8_t* t, int n) {   /* execute(), Pattern.cpp:64-66 */
    gen_t g; g.t = t; g.n = n; g.offset = -1; g.state = 0; g.tb = tb;
    return g;
}
/* operator++, Pattern.cpp:33-56.  NB `continue` inside do/while jumps to the condition. */
static void gen_next(gen_t* g) {
    const orc_table* tb = g->tb;
    do {
        if (g->n == 0) { g->state = 0; break; }
        int code = g->t[0];
        if (g->state == tb->inv[code]) {
            while (g->n != 0 && g->t[0] == code) { ++g->offset; ++g->t; --g->n; }
            continue;
        }
        int next = tb->base[g->state] + code;
        if (tb->check[next] == g->state) {
            g->state = next;
        } else if (g->state != 0) {
            g->state = tb->fail[g->state];
            continue;
        }
        ++g->offset; ++g->t; --g->n;
    } while (tb->check[tb->base[g->state]] != g->state);
}
static void gen_begin(gen_t* g) { if (g->state == 0) gen_next(g); }     /* begin(), Pattern.h:66 */
static int gen_at_end(const gen_t* g) { return g->n == 0 && g->state == 0; }   /* operator!=, Pattern.h:71-73 */
static int gen_pattern(const gen_t* g) {                                  /* operator*, Pattern.cpp:59-62 */
    int leaf = g->tb->base[g->state];
    return -g->tb->base[leaf];
}
/* HasCovered, Pattern.cpp:22-25 (the unsigned wrap-around makes it a two-sided test) */
static int has_covered(int pat_len, int offset, int pose) { return offset >= pose && offset - pose < pat_len; }

int orc_scan(const orc_table* tb, const uint8_t* codes, int n, int32_t* pids, int32_t* offsets, int cap) {
    gen_t g = gen_make(tb, codes, n);
    int count = 0;
    for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) {
        if (count < cap) { pids[count] = gen_pattern(&g); offsets[count] = g.offset; }
        ++count;
    }
    return count;
}
long orc_scan_many(const orc_table* tb, const uint8_t* codes, const int64_t* starts, int n_strings,
                   int32_t* pids, int32_t* offsets, int32_t* counts, long cap) {
    long total = 0;
    for (int s = 0; s < n_strings; ++s) {
        gen_t g = gen_make(tb, codes + starts[s], (int)(starts[s + 1] - starts[s]));
        int c = 0;
        for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) {
            if (total < cap) { pids[total] = gen_pattern(&g); offsets[total] = g.offset; }
            ++total; ++c;
        }
        counts[s] = c;
    }
    return total;
}

/* ======================================================================================
 * Board                          include/Game.h:59-151, src/Game.cpp
 * ==================================================================================== */
typedef struct {
    int cur, winner;
    uint8_t states[3][ORC_CELLS];        /* [player + 1][cell], Game.h:146 */
    int counts[3];
    int16_t record[ORC_CELLS + 1]; int nrec;
} board_t;

static void board_reset(board_t* b) {                /* Game.cpp:138-146 */
    memset(b, 0, sizeof *b);
    memset(b->states[P_NONE + 1], 1, ORC_CELLS);
    b->counts[P_NONE + 1] = ORC_CELLS;
    b->cur = P_BLACK; b->winner = P_NONE;
}
static int board_check_move(const board_t* b, int move) {   /* Game.cpp:80-82 */
    return move >= 0 && move < ORC_CELLS && b->states[P_NONE + 1][move];
}
static int board_check_end(board_t* b) {             /* Game.cpp:88-136 */
    if (b->cur == P_NONE) return 1;
    if (b->nrec == 0) return 0;
    int last = b->record[b->nrec - 1], cx = last % ORC_W, cy = last / ORC_W, lp = -b->cur;
    static const int SDX[4] = { 1, 0, 1, 1 }, SDY[4] = { 0, 1, -1, 1 };   /* :125 */
    int won = 0;
    for (int d = 0; d < 4 && !won; ++d) {
        int renju = 1;
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
            int x = cx, y = cy;
            for (int i = 1; i <= 5; ++i) {
                x += sgn * SDX[d]; y += sgn * SDY[d];
                if (x >= 0 && x < ORC_W && y >= 0 && y < ORC_H && b->states[lp + 1][y * ORC_W + x]) ++renju;
                else break;
            }
        }
        won = renju >= 5;
    }
    if (won) { b->winner = lp; b->cur = P_NONE; return 1; }
    if (b->counts[P_NONE + 1] == 0) { b->winner = P_NONE; b->cur = P_NONE; return 1; }
    return 0;
}
static int board_apply(board_t* b, int move, int check_victory) {   /* Game.cpp:37-47 */
    if (b->cur != P_NONE && board_check_move(b, move)) {
        b->states[b->cur + 1][move] = 1; b->counts[b->cur + 1] += 1;
        b->states[P_NONE + 1][move] = 0; b->counts[P_NONE + 1] -= 1;
        b->record[b->nrec++] = (int16_t)move;
        b->cur = -b->cur;
        if (check_victory) board_check_end(b);
    }
    return b->cur;
}
static int board_revert(board_t* b, int count) {     /* Game.cpp:49-62 */
    if (b->cur == P_NONE && count != 0) {
        b->cur = b->counts[P_BLACK + 1] == b->counts[P_WHITE + 1] ? P_BLACK : P_WHITE;
        b->winner = P_NONE;
    }
    for (int i = 0; b->nrec > 0 && i < count; ++i) {
        int mv = b->record[b->nrec - 1];
        b->states[-b->cur + 1][mv] = 0; b->counts[-b->cur + 1] -= 1;
        b->states[P_NONE + 1][mv] = 1; b->counts[P_NONE + 1] += 1;
        --b->nrec;
        b->cur = -b->cur;
    }
    return b->cur;
}

int orc_board_play(const int16_t* moves, int n_moves, int* out3) {
    board_t b; board_reset(&b);
    int applied = 0;
    for (int i = 0; i < n_moves; ++i) {
        int before = b.cur;
        if (board_apply(&b, moves[i], 1) != before) ++applied;
    }
    out3[0] = b.cur; out3[1] = b.winner; out3[2] = applied;
    return 0;
}

/* ======================================================================================
 * BoardMap                       include/Mapping.h:54-72, src/Mapping.cpp:11-77
 * ==================================================================================== */
typedef struct { board_t board; uint8_t line[ORC_LINES][28]; int len[ORC_LINES]; } bmap_t;

static void parse_index(int pose, int dir, int* index, int* offset) {   /* Mapping.cpp:11-25 */
    int x = pose % ORC_W, y = pose / ORC_W, off = ORC_MAX_PATTERN_LEN - 1;
    switch (dir) {
        case D_H:  *index = y; *offset = off + x; break;
        case D_V:  *index = ORC_H + x; *offset = off + y; break;
        case D_LD: *index = ORC_W + 2 * ORC_H - 1 + x - y; *offset = off + (x < y ? x : y); break;
        default:   *index = 2 * (ORC_W + ORC_H) - 1 + x + y;
                   *offset = off + (ORC_W - 1 - x < y ? ORC_W - 1 - x : y); break;
    }
}
static void bmap_reset(bmap_t* m) {                  /* Mapping.cpp:61-77 (hash omitted: no reader) */
    board_reset(&m->board);
    for (int l = 0; l < ORC_LINES; ++l) { memset(m->line[l], 3, 6); m->len[l] = 6; }
    for (int i = 0; i < ORC_CELLS; ++i)
        for (int d = 0; d < 4; ++d) { int idx, off; parse_index(i, d, &idx, &off); m->line[idx][m->len[idx]++] = 4; }
    for (int l = 0; l < ORC_LINES; ++l) { memset(m->line[l] + m->len[l], 3, 6); m->len[l] += 6; }
}
static const uint8_t* bmap_view(const bmap_t* m, int pose, int dir) {   /* lineView, Mapping.cpp:31-34 */
    int idx, off; parse_index(pose, dir, &idx, &off);
    return &m->line[idx][off - ORC_TARGET_LEN / 2];
}
static int bmap_apply(bmap_t* m, int move) {         /* Mapping.cpp:37-45 */
    for (int d = 0; d < 4; ++d) {
        int idx, off; parse_index(move, d, &idx, &off);
        m->line[idx][off] = (uint8_t)(m->board.cur == P_BLACK ? 1 : 2);
    }
    return board_apply(&m->board, move, 0);
}
static int bmap_revert(bmap_t* m, int count) {       /* Mapping.cpp:47-59 */
    for (int i = 0; i < count; ++i) {
        int move = m->board.record[m->board.nrec - 1];
        for (int d = 0; d < 4; ++d) { int idx, off; parse_index(move, d, &idx, &off); m->line[idx][off] = 4; }
        board_revert(&m->board, 1);
    }
    return m->board.cur;
}

int orc_line_view(const int16_t* moves, int n_moves, int pose, int dir, uint8_t* out13) {
    bmap_t* m = (bmap_t*)malloc(sizeof *m);
    bmap_reset(m);
    for (int i = 0; i < n_moves; ++i) bmap_apply(m, moves[i]);
    memcpy(out13, bmap_view(m, pose, dir), ORC_TARGET_LEN);
    free(m);
    return 0;
}
int orc_line_map(const int16_t* moves, int n_moves, uint8_t* out, int* lens) {
    bmap_t* m = (bmap_t*)malloc(sizeof *m);
    bmap_reset(m);
    for (int i = 0; i < n_moves; ++i) bmap_apply(m, moves[i]);
    int k = 0;
    for (int l = 0; l < ORC_LINES; ++l) { lens[l] = m->len[l]; memcpy(out + k, m->line[l], (size_t)m->len[l]); k += m->len[l]; }
    free(m);
    return k;
}

/* ======================================================================================
 * Evaluator                      include/Pattern.h:142-220, src/Pattern.cpp:76-416
 * ==================================================================================== */
enum { DIST_P = 8, DIST_C = 3, N_DIST = (ORC_CELLS + 1) * DIST_P + (ORC_CELLS + 1) * DIST_C };

struct orc_evaluator {
    bmap_t map;
    /* m_patternDist followed by m_compoundDist (Pattern.h:216-217), kept in ONE flat array in
     * declaration order so that the out-of-range index Compound::type == -1 produced by
     * Compound::locate (Pattern.cpp:484) aliases the same word as in the reference. */
    uint32_t dist[N_DIST];
    int density[2][2][ORC_CELLS];        /* [Group(player)][count|weight], Pattern.h:218 */
    int scores[4][ORC_CELLS];            /* Group(favour, perspective), Pattern.h:159-161,219 */
};
static long g_degenerate = 0;
long orc_degenerate_compounds(void) { return g_degenerate; }

static int grp1(int player) { return player == P_BLACK; }                              /* Pattern.h:154-156 */
static int grp2(int favour, int persp) { return ((favour == P_BLACK) << 1) | (persp == P_BLACK); }
static uint32_t* pdist(orc_evaluator* ev, int cell, int type) { return &ev->dist[cell * DIST_P + type]; }
static uint32_t* cdist(orc_evaluator* ev, int cell, int type) {
    return &ev->dist[(ORC_CELLS + 1) * DIST_P + cell * DIST_C + type];   /* type may be -1, see above */
}
/* Record::set / get, Pattern.cpp:390-416 */
static void rec_set_total(uint32_t* f, int delta, int player) { *f += (uint32_t)delta << (16 * grp1(player)); }
static void rec_set_flag(uint32_t* f, int delta, int favour, int persp, int dir) {
    unsigned offset = (unsigned)(4 * grp2(favour, persp) + dir) * 2;
    unsigned lower = 1u << offset, higher = lower << 1, mask = higher | lower;
    unsigned value = delta == 1 ? ((*f << 1) | lower) : ((*f >> 1) & ~higher);
    *f = (*f & ~mask) | (value & mask);
}
static unsigned rec_get_dir(uint32_t f, int favour, int persp, int dir) {
    return (f >> ((4 * grp2(favour, persp) + dir) * 2)) & 3u;
}
static unsigned rec_get_group(uint32_t f, int favour, int persp) { return (f >> (8 * grp2(favour, persp))) & 0xffu; }
static unsigned rec_get_total(uint32_t f, int player) { return (f >> (16 * grp1(player))) & 0xffffu; }

static void ev_reset(orc_evaluator* ev) {            /* Pattern.cpp:371-386 */
    bmap_reset(&ev->map);
    memset(ev->dist, 0, sizeof ev->dist);
    memset(ev->density, 0, sizeof ev->density);
    memset(ev->scores, 0, sizeof ev->scores);
}

/* ---- Compound, Pattern.h:100-139, Pattern.cpp:418-550 ---- */
typedef struct {
    int position, favour;
    int comp_dir[8], comp_type[8], ncomp;
    int type;
    gen_t gen; int gen_dir;
    int count, l3_count, triple_cross;
} compound_t;

static const int COMP_TYPES[3] = { T_LIVE3, T_DEAD3, T_LIVE2 };   /* Pattern.cpp:420-422 */

static int compound_test(orc_evaluator* ev, int pose, int player) {   /* Pattern.cpp:424-433 */
    unsigned bits = 0;
    for (int k = 0; k < 3; ++k) bits |= rec_get_group(*pdist(ev, pose, COMP_TYPES[k]), player, player);
    return (bits & (bits - 1)) != 0;
}
static void compound_locate(orc_evaluator* ev, compound_t* c) {        /* Pattern.cpp:440-486 */
    enum { S0, L2, LD3, To33, To43, To44 };
    int state = S0;
    for (int dir = 0; dir < 4; ++dir) {
        int cond = S0, count = 0;
        for (int k = 0; k < 3; ++k) {
            int ct = COMP_TYPES[k];
            switch (rec_get_dir(*pdist(ev, c->position, ct), c->favour, c->favour, dir)) {
                case 0: count = 0; break;
                case 1: count = 1; break;
                case 3: count = 2; break;
                default: break;           /* 0b10 cannot be produced by Record::set; count keeps its value */
            }
            if (count != 0) {
                if (ct == T_LIVE3) { c->l3_count += 1; cond = LD3; }
                else if (ct == T_DEAD3) cond = LD3;
                else cond = L2;
                for (int i = 0; i < count; ++i) { c->comp_dir[c->ncomp] = dir; c->comp_type[c->ncomp] = ct; ++c->ncomp; }
                break;
            }
        }
        if (cond == S0) continue;
        for (int i = 0; i < count; ++i) {
            int offset;
            if (state == S0) offset = 0;
            else if (state == L2 || state == LD3) offset = 1;
            else { c->triple_cross = 1; offset = state == To44 ? -cond : -1; }
            state = state + cond + offset;
        }
    }
    c->type = state - To33;
    if (c->type < 0) ++g_degenerate;
    unsigned bits = rec_get_group(*cdist(ev, c->position, c->type), c->favour, c->favour);
    c->count = __builtin_popcount(bits);
}
static void compound_update_pose(orc_evaluator* ev, compound_t* c, int delta, int pose, int comp_dir, int persp) {
    rec_set_flag(cdist(ev, pose, c->type), delta, c->favour, persp, comp_dir);   /* Pattern.cpp:545-550 */
    ev->scores[grp2(c->favour, persp)][pose] += delta * COMPOUND_SCORE;
}
static void compound_update_antis(orc_evaluator* ev, const orc_table* tb, compound_t* c, int delta, int k) {
    int comp_dir = c->comp_dir[k], comp_type = c->comp_type[k];            /* Pattern.cpp:520-543 */
    if (c->gen_dir != comp_dir) {
        c->gen = gen_make(tb, bmap_view(&ev->map, c->position, comp_dir), ORC_TARGET_LEN);
        c->gen_dir = comp_dir;
    }
    gen_begin(&c->gen);                   /* begin() advances the MEMBER when state == 0 (Pattern.h:66) ... */
    gen_t it = c->gen;                    /* ... and the loop then runs on a copy */
    for (; !gen_at_end(&it); gen_next(&it)) {
        const pat_t* p = &tb->pats[gen_pattern(&it)];
        int offset = it.offset;
        if (p->type == comp_type && has_covered(p->len, offset, ORC_TARGET_LEN / 2)
            && p->str[p->len - 1 - (offset - ORC_TARGET_LEN / 2)] == '_') {
            int current = shift(c->position, offset - ORC_TARGET_LEN / 2, comp_dir);
            for (int i = 0; i < p->len; ++i, current = shift(current, -1, comp_dir)) {
                char piece = p->str[p->len - 1 - i];
                if ((piece == '_' || piece == '^') && current != c->position)
                    compound_update_pose(ev, c, delta, current, comp_dir, -c->favour);
            }
            break;
        }
    }
}
static void compound_update(orc_evaluator* ev, const orc_table* tb, compound_t* c, int delta) {   /* Pattern.cpp:488-513 */
    for (int k = 0; k < c->ncomp; ++k) {
        if (2 * c->count + delta == -1) return;
        compound_update_pose(ev, c, delta, c->position, c->comp_dir[k], c->favour);    /* updateCritical :515-518 */
        compound_update_pose(ev, c, delta, c->position, c->comp_dir[k], -c->favour);
        if (!c->triple_cross && c->l3_count == 0) compound_update_antis(ev, tb, c, delta, k);
        if (2 * c->count + delta == 3)
            rec_set_total(cdist(ev, ORC_CELLS, c->type), delta, c->favour);
        c->count += delta;
    }
}

/* ---- Updater, Pattern.h:192-212, Pattern.cpp:111-302 ---- */
typedef struct { int pid, offset; } entry_t;
typedef struct {
    int delta, move, player;
    entry_t results[2][4][32]; int nres[2][4];
    int key_pose[128], key_player[128]; int ncomp;
} updater_t;

static void upd_match(orc_evaluator* ev, const orc_table* tb, updater_t* u, int dir) {   /* matchPatterns :128-136 */
    int slot = u->delta == 1;
    u->nres[slot][dir] = 0;
    gen_t g = gen_make(tb, bmap_view(&ev->map, u->move, dir), ORC_TARGET_LEN);
    for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) {
        int pid = gen_pattern(&g);
        if (has_covered(tb->pats[pid].len, g.offset, ORC_TARGET_LEN / 2)) {
            entry_t e; e.pid = pid; e.offset = g.offset;
            u->results[slot][dir][u->nres[slot][dir]++] = e;
        }
    }
}
static void upd_patterns(orc_evaluator* ev, const orc_table* tb, updater_t* u, int dir) {   /* updatePatterns :138-165 */
    int slot = u->delta == 1;
    for (int r = 0; r < u->nres[slot][dir]; ++r) {
        const pat_t* p = &tb->pats[u->results[slot][dir][r].pid];
        int offset = u->results[slot][dir][r].offset;
        if (p->type == T_FIVE) {
            ev->map.board.cur = P_NONE;
            ev->map.board.winner = p->favour;
            continue;
        }
        int current = shift(u->move, offset - ORC_TARGET_LEN / 2, dir);
        rec_set_total(pdist(ev, ORC_CELLS, p->type), u->delta, p->favour);
        for (int i = 0; i < p->len; ++i, current = shift(current, -1, dir)) {
            char piece = p->str[p->len - 1 - i];
            if (piece == '_' || piece == '^') {
                double multiplier = (dir == D_LD || dir == D_RD) ? 1.2 : 1;
                int score = (int)(u->delta * multiplier * p->score);
                if (piece == '_') {                  /* switch fall-through, :158-161 */
                    rec_set_flag(pdist(ev, current, p->type), u->delta, p->favour, p->favour, dir);
                    ev->scores[grp2(p->favour, p->favour)][current] += score;
                }
                rec_set_flag(pdist(ev, current, p->type), u->delta, p->favour, -p->favour, dir);
                ev->scores[grp2(p->favour, -p->favour)][current] += score;
            }
        }
    }
}
static void upd_compound(orc_evaluator* ev, const orc_table* tb, updater_t* u, int dir) {   /* updateCompound :167-197 */
    const uint8_t* view = bmap_view(&ev->map, u->move, dir);
    static const int players[2] = { P_WHITE, P_BLACK };
    for (int pi = 0; pi < 2; ++pi) {
        int player = players[pi], current = -1, offset = 0;
        for (int i = 0; i < ORC_TARGET_LEN; ++i) {
            if (view[i] != 4) continue;
            else if (current == -1) current = shift(u->move, i - ORC_TARGET_LEN / 2, dir);
            else current = shift(current, i - offset, dir);
            offset = i;
            if (ev->density[grp1(player)][0][current] < 2) continue;
            int found = 0;
            for (int k = 0; k < u->ncomp; ++k)
                if (u->key_pose[k] == current && u->key_player[k] == player) { found = 1; break; }
            if (found) continue;
            if (compound_test(ev, current, player)) {
                compound_t c; memset(&c, 0, sizeof c);
                c.position = current; c.favour = player; c.gen_dir = -1;
                compound_locate(ev, &c);
                if (u->ncomp < 128) { u->key_pose[u->ncomp] = current; u->key_player[u->ncomp] = player; ++u->ncomp; }
                compound_update(ev, tb, &c, u->delta);
            }
        }
    }
}
static void upd_block(orc_evaluator* ev, updater_t* u, int delta, int src) {   /* updateBlock :236-272, BlockView :94-109 */
    int mx = u->move % ORC_W, my = u->move / ORC_W;
    int x0 = mx - 3 < 0 ? 0 : mx - 3, x1 = mx + 3 > ORC_W - 1 ? ORC_W - 1 : mx + 3;
    int y0 = my - 3 < 0 ? 0 : my - 3, y1 = my + 3 > ORC_H - 1 ? ORC_H - 1 : my + 3;
    int* cnt = ev->density[grp1(src)][0];
    int* wgt = ev->density[grp1(src)][1];
    int* sc = ev->scores[grp2(src, src)];
    int mask_old[7][7];
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) mask_old[y - my + 3][x - mx + 3] = wgt[y * ORC_W + x] > 0;
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            int c = y * ORC_W + x, w = BLOCK_W[y - my + 3][x - mx + 3];
            wgt[c] += (wgt[c] < 0 ? -1 : 1) * delta * w;
            cnt[c] += (wgt[c] < 0 ? -1 : 1) * delta * (w > 0);      /* lazy sign_block reads the updated weight, :245,250 */
        }
    for (int pg = 1; pg >= 0; --pg)                                 /* { Black, White }, :253-265 */
        for (int cw = 0; cw < 2; ++cw) {
            int* v = &ev->density[pg][cw][u->move];
            if (delta == 1) { *v *= -1; *v -= 1; } else { *v += 1; *v *= -1; }
        }
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            int c = y * ORC_W + x;
            sc[c] += BLOCK_SCORE * ((wgt[c] > 0) - mask_old[y - my + 3][x - mx + 3]);
        }
    int count = ev->density[grp1(-src)][0][u->move];
    if (count != 0 && count != -1) ev->scores[grp2(-src, -src)][u->move] -= delta * BLOCK_SCORE;
}
static void upd_phase_reset(updater_t* u, int delta, int move, int player) {   /* :111-117 */
    u->delta = delta; u->move = move; u->player = player; u->ncomp = 0;
}
static void upd_move(orc_evaluator* ev, const orc_table* tb, updater_t* u, int move, int src) {   /* updateMove :274-302 */
    upd_phase_reset(u, -1, move, src);
    for (int d = 0; d < 4; ++d) upd_match(ev, tb, u, d);
    for (int d = 0; d < 4; ++d) upd_compound(ev, tb, u, d);
    for (int d = 0; d < 4; ++d) upd_patterns(ev, tb, u, d);
    if (src != P_NONE) {
        bmap_apply(&ev->map, move);
        upd_block(ev, u, 1, src);
    } else {
        bmap_revert(&ev->map, 1);
        upd_block(ev, u, -1, ev->map.board.cur);
    }
    upd_phase_reset(u, 1, move, src);
    for (int d = 0; d < 4; ++d) upd_match(ev, tb, u, d);
    for (int d = 0; d < 4; ++d) upd_patterns(ev, tb, u, d);
    for (int d = 0; d < 4; ++d) upd_compound(ev, tb, u, d);
}

/* Evaluator::applyMove, Pattern.cpp:310-335.  Returns 1 where the reference's self-check
 * would throw (occupied cell with a non-zero score, or a negative score). */
static int ev_apply(orc_evaluator* ev, const orc_table* tb, updater_t* u, int move) {
    board_t* b = &ev->map.board;
    if (b->cur != P_NONE && board_check_move(b, move)) upd_move(ev, tb, u, move, b->cur);
    for (int i = 0; i < ORC_CELLS; ++i)
        for (int j = 0; j < 4; ++j) {
            if (!b->states[P_NONE + 1][i]) { if (ev->scores[j][i] != 0) return 1; }
            else if (ev->scores[j][i] < 0) return 1;
        }
    return 0;
}
/* Evaluator::revertMove, Pattern.cpp:337-342 */
static void ev_revert(orc_evaluator* ev, const orc_table* tb, updater_t* u, int count) {
    board_t* b = &ev->map.board;
    for (int i = 0; i < count && b->nrec > 0; ++i) upd_move(ev, tb, u, b->record[b->nrec - 1], P_NONE);
}

static orc_evaluator* g_ev = NULL;
static updater_t* g_upd = NULL;
static void ev_global(void) {
    if (!g_ev) { g_ev = (orc_evaluator*)malloc(sizeof *g_ev); g_upd = (updater_t*)malloc(sizeof *g_upd); }
}
static void ev_dump(orc_evaluator* ev, int32_t* scores, uint16_t* pat_totals, uint16_t* cmp_totals,
                    int8_t* winner, int8_t* cur_player) {
    if (scores) for (int g = 0; g < 4; ++g) for (int i = 0; i < ORC_CELLS; ++i) scores[g * ORC_CELLS + i] = ev->scores[g][i];
    if (pat_totals) for (int t = 0; t < 8; ++t) {
        pat_totals[t] = (uint16_t)rec_get_total(*pdist(ev, ORC_CELLS, t), P_WHITE);
        pat_totals[8 + t] = (uint16_t)rec_get_total(*pdist(ev, ORC_CELLS, t), P_BLACK);
    }
    if (cmp_totals) for (int t = 0; t < 3; ++t) {
        cmp_totals[t] = (uint16_t)rec_get_total(*cdist(ev, ORC_CELLS, t), P_WHITE);
        cmp_totals[3 + t] = (uint16_t)rec_get_total(*cdist(ev, ORC_CELLS, t), P_BLACK);
    }
    if (winner) *winner = (int8_t)ev->map.board.winner;
    if (cur_player) *cur_player = (int8_t)ev->map.board.cur;
}

int orc_eval_moves(const int16_t* moves, int n_moves, int32_t* scores, uint16_t* pat_totals,
                   uint16_t* cmp_totals, int8_t* winner, int8_t* cur_player) {
    const orc_table* tb = orc_table_default();
    ev_global();
    ev_reset(g_ev);
    for (int i = 0; i < n_moves; ++i)
        if (ev_apply(g_ev, tb, g_upd, moves[i])) { ev_reset(g_ev); return 1; }
    ev_dump(g_ev, scores, pat_totals, cmp_totals, winner, cur_player);
    return 0;
}
int orc_eval_batch(const int16_t* moves, const int64_t* starts, int n_pos, int32_t* scores,
                   uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner, int8_t* cur_player) {
    int bad = 0;
    for (int p = 0; p < n_pos; ++p)
        bad += orc_eval_moves(moves + starts[p], (int)(starts[p + 1] - starts[p]),
                              scores ? scores + (size_t)p * 4 * ORC_CELLS : NULL,
                              pat_totals ? pat_totals + (size_t)p * 16 : NULL,
                              cmp_totals ? cmp_totals + (size_t)p * 6 : NULL,
                              winner ? winner + p : NULL, cur_player ? cur_player + p : NULL);
    return bad;
}
int orc_eval_flags(uint32_t* pattern_flags, uint32_t* compound_flags, int32_t* density) {
    ev_global();
    for (int c = 0; c < ORC_CELLS; ++c) {
        for (int t = 0; t < 8; ++t) pattern_flags[c * 8 + t] = *pdist(g_ev, c, t);
        for (int t = 0; t < 3; ++t) compound_flags[c * 3 + t] = *cdist(g_ev, c, t);
    }
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < ORC_CELLS; ++c)
        density[(a * 2 + b) * ORC_CELLS + c] = g_ev->density[a][b][c];
    return 0;
}
/* revert is restated for completeness (Evaluator::syncWithBoard uses it) and exercised by
 * the apply/revert symmetry test. */
int orc_eval_apply_revert(const int16_t* moves, int n_moves, int n_revert, int32_t* scores,
                          uint16_t* pat_totals, uint16_t* cmp_totals) {
    const orc_table* tb = orc_table_default();
    ev_global();
    ev_reset(g_ev);
    for (int i = 0; i < n_moves; ++i) if (ev_apply(g_ev, tb, g_upd, moves[i])) return 1;
    ev_revert(g_ev, tb, g_upd, n_revert);
    ev_dump(g_ev, scores, pat_totals, cmp_totals, NULL, NULL);
    return 0;
}

long orc_linescan_batch(const int16_t* moves, const int64_t* starts, int n_pos) {
    const orc_table* tb = orc_table_default();
    bmap_t* m = (bmap_t*)malloc(sizeof *m);
    long total = 0;
    for (int p = 0; p < n_pos; ++p) {
        bmap_reset(m);
        for (int64_t i = starts[p]; i < starts[p + 1]; ++i) bmap_apply(m, moves[i]);
        for (int l = 0; l < ORC_LINES; ++l) {
            gen_t g = gen_make(tb, m->line[l], m->len[l]);
            for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) ++total;
        }
    }
    free(m);
    return total;
}

/* ======================================================================================
 * Rollout                        include/algorithms/MonteCarlo.hpp:37-47, src/Game.cpp:64-73
 * ==================================================================================== */
static void rollout_setup(board_t* b, const int16_t* moves, int n_moves) {
    board_reset(b);
    for (int i = 0; i < n_moves; ++i) board_apply(b, moves[i], 0);   /* Policy::applyMove, MCTS.cpp:48-50 */
    board_check_end(b);                                              /* MCTS.cpp:166 */
}
/* getRandomMove's probe (Game.cpp:68-71) with r standing in for rnd(rnd_eng) */
static int probe_move(const board_t* b, int r) {
    int id = r;
    while (!b->states[P_NONE + 1][id]) id = (id + 1) % ORC_CELLS;
    return id;
}
int orc_rollout_injected(const int16_t* moves, int n_moves, const uint8_t* r_stream, int stream_len, int* n_played) {
    board_t b; rollout_setup(&b, moves, n_moves);
    int total = 0;
    for (int result = b.cur; result != P_NONE; ++total) {            /* RandomRollout, MonteCarlo.hpp:39-41 */
        if (total >= stream_len) { *n_played = total; return -2; }
        result = board_apply(&b, probe_move(&b, r_stream[total]), 1);
    }
    *n_played = total;
    return b.winner;
}

/* Philox4x32-10 (Salmon et al., SC'11): the counter-based generator both sides share for
 * the injected-stream parity protocol.  Not part of the reference. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
int orc_rollout_philox(const int16_t* moves, int n_moves, uint64_t key, uint32_t position,
                       uint32_t rollout, uint32_t ctr_hi, int* n_played) {
    board_t b; rollout_setup(&b, moves, n_moves);
    uint32_t k[2] = { (uint32_t)key, (uint32_t)(key >> 32) }, w[4] = { 0, 0, 0, 0 };
    int total = 0;
    for (int result = b.cur; result != P_NONE; ++total) {
        if ((total & 3) == 0) {
            uint32_t ctr[4] = { (uint32_t)(total >> 2), rollout, position, ctr_hi };
            orc_philox4x32_10(ctr, k, w);
        }
        int r = (int)(((uint64_t)w[total & 3] * ORC_CELLS) >> 32);
        result = board_apply(&b, probe_move(&b, r), 1);
    }
    *n_played = total;
    return b.winner;
}
int orc_rollout_philox_batch(const int16_t* moves, const int64_t* starts, int n_pos, int rollouts_per_pos,
                             uint64_t key, uint32_t ctr_hi, int pos_base, int8_t* winners, uint8_t* lengths,
                             int32_t* wdb) {
    for (int p = 0; p < n_pos; ++p) {
        if (wdb) { wdb[p * 3] = wdb[p * 3 + 1] = wdb[p * 3 + 2] = 0; }
        for (int r = 0; r < rollouts_per_pos; ++r) {
            int len = 0;
            int w = orc_rollout_philox(moves + starts[p], (int)(starts[p + 1] - starts[p]), key,
                                       (uint32_t)(pos_base + p), (uint32_t)r, ctr_hi, &len);
            if (winners) winners[(size_t)p * rollouts_per_pos + r] = (int8_t)w;
            if (lengths) lengths[(size_t)p * rollouts_per_pos + r] = (uint8_t)len;
            if (wdb) wdb[p * 3 + w + 1] += 1;
        }
    }
    return 0;
}

/* ======================================================================================
 * From-scratch model (NOT how the reference computes; how the CUDA kernel does)
 * --------------------------------------------------------------------------------------
 * Recomputes m_scores / totals from the final position alone, following SURVEY.md
 * Appendix A.2-A.4: scan every line once, add pattern scores and per-cell flags, add the
 * density block score, then detect compounds from the flags.  It exists so that the
 * equivalence "incremental replay == from-scratch" can be checked on the CPU at scale
 * (tests/test_oracle_scratch.py) and so that a failing kernel can be bisected stage by
 * stage.  lead/trail = number of '?' pads per line (the reference stores 6/6; the kernel
 * uses the minimal 1/2); min_len = shortest line scanned (1 = all 88 lines, 5 = the 72
 * lines that can hold a pattern).
 * ==================================================================================== */
static long g_scratch_gate_only = 0;    /* Test passed but the density gate (< 2) blocked it */
long orc_scratch_gate_blocks(void) { return g_scratch_gate_only; }

typedef struct { int cell0, stride, len, dir; } line_t;
static int scratch_lines(line_t* out, int min_len) {
    int n = 0;
    for (int y = 0; y < ORC_H; ++y) { line_t l = { y * ORC_W, 1, ORC_W, D_H }; if (l.len >= min_len) out[n++] = l; }
    for (int x = 0; x < ORC_W; ++x) { line_t l = { x, ORC_W, ORC_H, D_V }; if (l.len >= min_len) out[n++] = l; }
    for (int k = -(ORC_H - 1); k < ORC_W; ++k) {          /* LeftDiag, x - y = k, runs (+1,+1) */
        int x0 = k > 0 ? k : 0, y0 = k > 0 ? 0 : -k, len = ORC_W - (k > 0 ? k : -k);
        line_t l = { y0 * ORC_W + x0, ORC_W + 1, len, D_LD }; if (l.len >= min_len) out[n++] = l;
    }
    for (int k = 0; k < ORC_W + ORC_H - 1; ++k) {         /* RightDiag, x + y = k, runs (-1,+1) */
        int x0 = k < ORC_W ? k : ORC_W - 1, y0 = k - x0, len = (k < ORC_W ? k : 2 * (ORC_W - 1) - k) + 1;
        line_t l = { y0 * ORC_W + x0, ORC_W - 1, len, D_RD }; if (l.len >= min_len) out[n++] = l;
    }
    return n;
}

int orc_eval_scratch(const uint8_t* cells, int lead, int trail, int min_len, int32_t* scores_out,
                     uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner) {
    const orc_table* tb = orc_table_default();
    static int scores[4][ORC_CELLS];
    static uint8_t flag[ORC_CELLS][2][3][4];             /* [cell][player grp][L3,D3,L2][dir], saturating at 2 */
    int ptot[2][8], ctot[2][3], win = 0;
    memset(scores, 0, sizeof scores); memset(flag, 0, sizeof flag);
    memset(ptot, 0, sizeof ptot); memset(ctot, 0, sizeof ctot);
    line_t lines[ORC_LINES];
    int nl = scratch_lines(lines, min_len);
    for (int li = 0; li < nl; ++li) {                    /* A.2 */
        uint8_t buf[64]; int n = 0;
        const line_t* L = &lines[li];
        for (int i = 0; i < lead; ++i) buf[n++] = 3;
        for (int i = 0; i < L->len; ++i) { int v = cells[L->cell0 + i * L->stride]; buf[n++] = (uint8_t)(v == 0 ? 4 : v); }
        for (int i = 0; i < trail; ++i) buf[n++] = 3;
        gen_t g = gen_make(tb, buf, n);
        for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) {
            const pat_t* p = &tb->pats[gen_pattern(&g)];
            if (p->type == T_FIVE) { win = p->favour; continue; }
            ptot[grp1(p->favour)][p->type] += 1;
            int v = (L->dir == D_LD || L->dir == D_RD) ? (int)(1.2 * p->score) : p->score;
            for (int i = 0; i < p->len; ++i) {
                char piece = p->str[p->len - 1 - i];
                if (piece != '_' && piece != '^') continue;
                int cell = L->cell0 + (g.offset - lead - i) * L->stride;
                if (piece == '_') {
                    scores[grp2(p->favour, p->favour)][cell] += v;
                    int k = p->type == T_LIVE3 ? 0 : p->type == T_DEAD3 ? 1 : p->type == T_LIVE2 ? 2 : -1;
                    if (k >= 0 && flag[cell][grp1(p->favour)][k][L->dir] < 2) flag[cell][grp1(p->favour)][k][L->dir] += 1;
                }
                scores[grp2(p->favour, -p->favour)][cell] += v;
            }
        }
    }
    static int dcount[2][ORC_CELLS];
    for (int c = 0; c < ORC_CELLS; ++c) {                 /* A.3 */
        if (cells[c]) continue;
        int cx = c % ORC_W, cy = c / ORC_W;
        for (int pg = 0; pg < 2; ++pg) {
            int stone = pg ? 1 : 2, cnt = 0, wsum = 0;
            for (int dy = -3; dy <= 3; ++dy) for (int dx = -3; dx <= 3; ++dx) {
                int x = cx + dx, y = cy + dy;
                if (x < 0 || x >= ORC_W || y < 0 || y >= ORC_H) continue;
                if (cells[y * ORC_W + x] == stone) { wsum += BLOCK_W[dy + 3][dx + 3]; cnt += BLOCK_W[dy + 3][dx + 3] > 0; }
            }
            dcount[pg][c] = cnt;
            if (wsum > 0) scores[pg ? 3 : 0][c] += BLOCK_SCORE;
        }
    }
    for (int c = 0; c < ORC_CELLS; ++c) {                 /* A.4 */
        if (cells[c]) continue;
        for (int pg = 0; pg < 2; ++pg) {
            int P = pg ? P_BLACK : P_WHITE;
            unsigned bits = 0;
            for (int k = 0; k < 3; ++k) for (int d = 0; d < 4; ++d) if (flag[c][pg][k][d]) bits |= (flag[c][pg][k][d] == 2 ? 3u : 1u) << (2 * d);
            if (!(bits & (bits - 1))) continue;
            if (dcount[pg][c] < 2) { ++g_scratch_gate_only; continue; }
            enum { S0, L2, LD3, To33, To43, To44 };
            int state = S0, l3 = 0, triple = 0, ncomp = 0, cdir[8], ctype[8];
            for (int d = 0; d < 4; ++d) {
                int cond = S0, count = 0;
                for (int k = 0; k < 3; ++k) {
                    count = flag[c][pg][k][d];
                    if (count) {
                        if (k == 0) { l3 += 1; cond = LD3; } else if (k == 1) cond = LD3; else cond = L2;
                        for (int i = 0; i < count; ++i) { cdir[ncomp] = d; ctype[ncomp] = COMP_TYPES[k]; ++ncomp; }
                        break;
                    }
                }
                if (cond == S0) continue;
                for (int i = 0; i < count; ++i) {
                    int off;
                    if (state == S0) off = 0; else if (state == L2 || state == LD3) off = 1;
                    else { triple = 1; off = state == To44 ? -cond : -1; }
                    state += cond + off;
                }
            }
            int type = state - To33;
            if (type < 0) return 2;                       /* degenerate: the model has no defined answer */
            for (int k = 0; k < ncomp; ++k) {
                scores[grp2(P, P)][c] += COMPOUND_SCORE;
                scores[grp2(P, -P)][c] += COMPOUND_SCORE;
                if (!triple && l3 == 0) {                 /* updateAntis on the 13-window around c */
                    uint8_t win13[ORC_TARGET_LEN];
                    int cx = c % ORC_W, cy = c / ORC_W, d = cdir[k];
                    for (int i = 0; i < ORC_TARGET_LEN; ++i) {
                        int x = cx + DIR_DX[d] * (i - 6), y = cy + DIR_DY[d] * (i - 6);
                        if (x < 0 || x >= ORC_W || y < 0 || y >= ORC_H) win13[i] = 3;
                        else { int v = cells[y * ORC_W + x]; win13[i] = (uint8_t)(v == 0 ? 4 : v); }
                    }
                    gen_t g = gen_make(tb, win13, ORC_TARGET_LEN);
                    for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) {
                        const pat_t* p = &tb->pats[gen_pattern(&g)];
                        if (p->type == ctype[k] && has_covered(p->len, g.offset, 6) && p->str[p->len - 1 - (g.offset - 6)] == '_') {
                            int cur = shift(c, g.offset - 6, d);
                            for (int i = 0; i < p->len; ++i, cur = shift(cur, -1, d)) {
                                char piece = p->str[p->len - 1 - i];
                                if ((piece == '_' || piece == '^') && cur != c) scores[grp2(P, -P)][cur] += COMPOUND_SCORE;
                            }
                            break;
                        }
                    }
                }
                if (k == 1) ctot[pg][type] += 1;
            }
        }
    }
    if (scores_out) memcpy(scores_out, scores, sizeof scores);
    if (pat_totals) for (int pg = 0; pg < 2; ++pg) for (int t = 0; t < 8; ++t) pat_totals[pg * 8 + t] = (uint16_t)ptot[pg][t];
    if (cmp_totals) for (int pg = 0; pg < 2; ++pg) for (int t = 0; t < 3; ++t) cmp_totals[pg * 3 + t] = (uint16_t)ctot[pg][t];
    if (winner) *winner = (int8_t)win;
    return 0;
}

/* Batch driver: positions given as move lists (black first, alternating); compares nothing,
 * just evaluates.  cells are derived by placing stones alternately. */
int orc_eval_scratch_batch(const int16_t* moves, const int64_t* starts, int n_pos, int lead, int trail, int min_len,
                           int32_t* scores, uint16_t* pat_totals, uint16_t* cmp_totals, int8_t* winner) {
    int bad = 0;
    for (int p = 0; p < n_pos; ++p) {
        uint8_t cells[ORC_CELLS]; memset(cells, 0, sizeof cells);
        int k = 0;
        for (int64_t i = starts[p]; i < starts[p + 1]; ++i, ++k) cells[moves[i]] = (uint8_t)((k & 1) ? 2 : 1);
        bad += orc_eval_scratch(cells, lead, trail, min_len, scores ? scores + (size_t)p * 4 * ORC_CELLS : NULL,
                                pat_totals ? pat_totals + (size_t)p * 16 : NULL,
                                cmp_totals ? cmp_totals + (size_t)p * 6 : NULL, winner ? winner + p : NULL) != 0;
    }
    return bad;
}

/* ======================================================================================
 * Exhaustive single-line checks behind "incremental replay == from-scratch evaluation"
 * --------------------------------------------------------------------------------------
 * The reference updates its state from 13-symbol windows around each move; the CUDA kernel
 * scans whole lines of the final position.  Because the matcher is not a textbook automaton
 * (no output links, emissions depend on up to two symbols of left context) the two are equal
 * only if, for EVERY line content:
 *   T4  the emissions covering a cell m found in the 13-window around m are exactly (same
 *       patterns, same positions, same order) those covering m in the scan of the whole line;
 *   T5  changing cell m does not change any whole-line emission that does not cover m (Five
 *       emissions excepted: where a run of five-or-more is reported moves with the run's length,
 *       but a Five only sets the winner and never contributes to scores or totals);
 * and the compound logic is well defined only if
 *   T1  no cell has, on one line and for one player, exactly one '_' pattern of a class and
 *       two or more of a lower-priority class (Compound::locate would yield type -1);
 *   T2  no cell has three or more '_' patterns of one class on one line (the 2-bit unary
 *       counters of Record::set would saturate and lose a later removal).
 * All lines of length 5..max_len over {x, o, blank} are enumerated with the reference's 6/6
 * padding.  out[0..3] = violations of T1, T2, T4, T5; out[4] = lines, out[5] = emissions,
 * out[7] = T5-style differences that involve only Five emissions (informational).
 * T3 (every L3/D3/L2 pattern has >= 2 own stones within distance 3 of each of its '_' cells,
 * which makes the density gate of updateCompound redundant) is a property of the table:
 * out[6] = patterns violating it.
 * ==================================================================================== */
typedef struct { int pid, end; } em_t;

static int scan_list(const orc_table* tb, const uint8_t* s, int n, em_t* out, int cap) {
    gen_t g = gen_make(tb, s, n);
    int c = 0;
    for (gen_begin(&g); !gen_at_end(&g); gen_next(&g)) {
        if (c < cap) { out[c].pid = gen_pattern(&g); out[c].end = g.offset; }
        ++c;
    }
    return c;
}
static int covers(const orc_table* tb, const em_t* e, int pos) { return e->end >= pos && e->end - pos < tb->pats[e->pid].len; }

int orc_line_theorems(int max_len, long* out) {
    const orc_table* tb = orc_table_default();
    memset(out, 0, sizeof(long) * 8);
    for (int p = 0; p < tb->n_pats; ++p) {                 /* T3 */
        const pat_t* pt = &tb->pats[p];
        if (pt->type != T_LIVE3 && pt->type != T_DEAD3 && pt->type != T_LIVE2) continue;
        char own = pt->favour == P_BLACK ? 'x' : 'o';
        for (int i = 0; i < pt->len; ++i) {
            if (pt->str[i] != '_') continue;
            int near = 0;
            for (int j = 0; j < pt->len; ++j) if (pt->str[j] == own && abs(i - j) <= 3) ++near;
            if (near < 2) { out[6] += 1; break; }
        }
    }
    for (int L = 5; L <= max_len; ++L) {
        long total = 1;
        for (int i = 0; i < L; ++i) total *= 3;
        for (long code = 0; code < total; ++code) {
            uint8_t line[40];
            int n = 0;
            long c = code;
            for (int i = 0; i < 6; ++i) line[n++] = 3;
            for (int i = 0; i < L; ++i) { int v = (int)(c % 3); c /= 3; line[n++] = (uint8_t)(v == 0 ? 4 : v); }
            for (int i = 0; i < 6; ++i) line[n++] = 3;
            em_t full[64];
            int nf = scan_list(tb, line, n, full, 64);
            out[4] += 1; out[5] += nf;
            /* T1 / T2: '_' counts per cell, player, class */
            int cnt[15][2][3];
            memset(cnt, 0, sizeof cnt);
            for (int e = 0; e < nf && e < 64; ++e) {
                const pat_t* pt = &tb->pats[full[e].pid];
                int k = pt->type == T_LIVE3 ? 0 : pt->type == T_DEAD3 ? 1 : pt->type == T_LIVE2 ? 2 : -1;
                if (k < 0) continue;
                for (int i = 0; i < pt->len; ++i)
                    if (pt->str[pt->len - 1 - i] == '_') cnt[full[e].end - i - 6][grp1(pt->favour)][k] += 1;
            }
            for (int m = 0; m < L; ++m) for (int pg = 0; pg < 2; ++pg) {
                for (int k = 0; k < 3; ++k) if (cnt[m][pg][k] >= 3) out[1] += 1;
                for (int k = 0; k < 3; ++k) {
                    if (cnt[m][pg][k] == 0) continue;
                    if (cnt[m][pg][k] == 1) for (int k2 = k + 1; k2 < 3; ++k2) if (cnt[m][pg][k2] >= 2) out[0] += 1;
                    break;
                }
            }
            for (int m = 0; m < L; ++m) {
                /* T4: window around m (padded index m + 6) vs whole line, emissions covering m, in order */
                em_t win[32];
                int nw = scan_list(tb, line + m, ORC_TARGET_LEN, win, 32);
                int a = 0, b = 0, bad = 0;
                for (;;) {
                    while (a < nw && !covers(tb, &win[a], 6)) ++a;
                    while (b < nf && !covers(tb, &full[b], m + 6)) ++b;
                    if (a >= nw || b >= nf) { bad = (a < nw) != (b < nf); break; }
                    if (win[a].pid != full[b].pid || win[a].end + m != full[b].end) { bad = 1; break; }
                    ++a; ++b;
                }
                out[2] += bad;
                /* T5: change cell m, emissions not covering m must not move */
                uint8_t keep = line[m + 6];
                for (int v = 1; v <= 4; ++v) {
                    if (v == 3 || v == keep) continue;
                    line[m + 6] = (uint8_t)v;
                    em_t alt[64];
                    int na = scan_list(tb, line, n, alt, 64);
                    for (int five = 0; five < 2; ++five) {           /* pass 0: all but Five, pass 1: Five only */
                        int i = 0, j = 0, diff = 0;
                        for (;;) {
                            while (i < nf && (covers(tb, &full[i], m + 6) || (tb->pats[full[i].pid].type == T_FIVE) != five)) ++i;
                            while (j < na && (covers(tb, &alt[j], m + 6) || (tb->pats[alt[j].pid].type == T_FIVE) != five)) ++j;
                            if (i >= nf || j >= na) { diff = (i < nf) != (j < na); break; }
                            if (full[i].pid != alt[j].pid || full[i].end != alt[j].end) { diff = 1; break; }
                            ++i; ++j;
                        }
                        out[five ? 7 : 3] += diff;
                    }
                }
                line[m + 6] = keep;
            }
        }
    }
    return 0;
}
