// ORACLE SCAFFOLDING -- test infrastructure, not product code.
// Replaces core/lib/src/utils/Persistence.cpp, which does not compile on Linux
// (std::ifstream(const wchar_t*) at Persistence.cpp:9,21,28 is MSVC-only).  The
// persistence file only stores Zobrist keys (Mapping.cpp:81-96) whose hash has
// no reader anywhere on the hot path, so returning "no data" is behaviour-neutral.
#include "utils/Persistence.h"
using namespace Gomoku;
json Persistence::Load(std::string_view) { return json(); }
void Persistence::Save(std::string_view, json) {}
