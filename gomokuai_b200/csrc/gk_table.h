// gk_table.h -- host-side table compiler: pattern prototypes -> flat transducer + scan tape.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gk_format.h"

namespace gk {

struct Proto { std::string text; int type; int score; };   // "+xxxxx" / "-_oooo_" as Pattern.cpp:14-18

struct PatternInfo { std::string str; int favour; int type; int score; };   // favour: +1 black, -1 white

struct HostTable {
    std::vector<PatternInfo> patterns;   // in the reference's order: pattern id == its index in PatternSearch::m_patterns
    int n_states = 0;                    // dense ids, 0 = root
    std::vector<uint32_t> trans;         // n_states * 4 transition words (gk_format.h)
    std::vector<int16_t> flush;          // per state: pattern id emitted if the input ends here, else -1
    int start_state = 0;                 // state after the single leading '?' of a board line
    int trail_pad = 0;                   // trailing '?' symbols after which no state can emit any more (>= 2)
    int tape_steps = 0;                  // scan steps per board (longest lane chain, even)
    // device encodings (gk_format.h)
    std::vector<uint16_t> dev_next;      // (n_clones + n_states) * 4 row offsets, indexed by raw cell value
    std::vector<uint32_t> dev_erec;      // n_clones emission records
    int n_clones = 0;
    int root_off = 0, start_off = 0;     // byte offsets of the root row / the line-start row in dev_next
    int list_cap = 0;                    // emission-list slots per lane: max emitting steps of any lane chain (+ rounding)
    std::vector<PatRec> patrec;
    std::vector<uint16_t> tape_src;      // tape_steps * 32
    std::vector<uint16_t> tape_info;     // tape_steps * 32
    int sync_depth = 0;                  // symbols after which the state no longer depends on the start state (0 = not synchronizing)
    std::string error;
};

// The reference's 41 prototypes (src/Pattern.cpp:554-596).
const std::vector<Proto>& default_protos();

// Returns false and fills out.error when the prototypes cannot be represented.
bool compile_table(const std::vector<Proto>& protos, HostTable& out);

}  // namespace gk
