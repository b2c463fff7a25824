// module.cpp -- pybind11 module `CorePyExt` with the reference's Python surface
// (core/py_ext/src/{module.cpp,game_ext.hpp,mcts_ext.hpp,policy_ext.hpp}): GameConfig, Player,
// Position, Board, Node, Policy, MCTS, RandomPolicy -- same names, arguments and defaults -- over
// the host mirror in game.h / mcts.h, whose simulate slots run on the GPU.  Vectors that the
// reference passes as Eigen::VectorXf cross as numpy float32[225].  New: RootParallelSearch.
#include <pybind11/chrono.h>
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../../../include/gomoku_b200.h"
#include "mcts.h"
#include "root_parallel.h"

namespace py = pybind11;
using namespace gomoku;
using namespace pybind11::literals;

static py::array_t<float> to_numpy(const Probs& p) { return py::array_t<float>(static_cast<py::ssize_t>(p.size()), p.data()); }

static std::string pos_str(const Position& p) { return "(" + std::to_string(p.x()) + ", " + std::to_string(p.y()) + ")"; }
static const char* player_str(Player p) { return p == Player::Black ? "Player Black" : p == Player::White ? "Player White" : "No Player"; }

PYBIND11_MODULE(CorePyExt, mod) {
    mod.doc() = "Gomoku AI core module (B200-native hot path)";

    mod.def("seed", &set_seed, "seed"_a, "Seed the playout streams and the Dirichlet-noise engine (the reference seeds from random_device).");
    mod.def("init", [](int device) { if (gk_init(device) != GK_OK) throw std::runtime_error(gk_last_error()); },
            "device"_a = 0, "Bind this process to one GPU (called implicitly by gomokuai_b200.core).");

    // ---- game_ext.hpp -------------------------------------------------------------------------------------
    mod.add_object("GameConfig", py::dict("width"_a = int(WIDTH), "height"_a = int(HEIGHT), "board_size"_a = int(BOARD_SIZE),
                                          "max_renju"_a = int(MAX_RENJU)));

    py::enum_<Player>(mod, "Player", "Gomoku player types")
        .value("white", Player::White)
        .value("none", Player::None)
        .value("black", Player::Black)
        .def("__float__", [](Player p) { return static_cast<double>(p); })
        .def("__neg__", [](Player p) { return -p; })
        .def_static("calc_score", [](Player a, Player b) { return double(CalcScore(a, b)); })
        .def_static("calc_score", [](Player a, double v) { return double(CalcScore(a, float(v))); });

    py::class_<Position>(mod, "Position", "Gomoku board positions")
        .def(py::init<int>())
        .def(py::init<int, int>())
        .def_readwrite("id", &Position::id)
        .def_property("x", &Position::x, [](Position& p, int x) { p.id = short(p.y() * WIDTH + x); })
        .def_property("y", &Position::y, [](Position& p, int y) { p.id = short(p.id + (y - p.y()) * WIDTH); })
        .def("__int__", [](const Position& p) { return int(p.id); })
        .def("__index__", [](const Position& p) { return int(p.id); })
        .def("__hash__", [](const Position& p) { return std::size_t(p.id); })
        .def("__eq__", [](const Position& a, const Position& b) { return a.id == b.id; })
        .def("__len__", [](const Position&) { return 2; })
        .def("__repr__", [](const Position& p) { return "Position" + pos_str(p); })
        .def("__str__", [](const Position& p) { return pos_str(p); })
        .def("__iter__", [](const Position& p) { return py::iter(py::make_tuple(p.x(), p.y())); });
    py::implicitly_convertible<int, Position>();

    py::class_<Board>(mod, "Board", "Gomoku game board")
        .def(py::init<>())
        .def("apply_move", &Board::applyMove, "move"_a, "checkVictory"_a = true)
        .def("revert_move", &Board::revertMove, "count"_a = 1)
        .def("random_move", &Board::getRandomMove)
        .def("check_move", &Board::checkMove)
        .def("check_end", &Board::checkGameEnd)
        .def("reset", &Board::reset)
        .def_readonly("move_record", &Board::m_moveRecord)
        .def_property_readonly("last_move", [](const Board& b) { return b.m_moveRecord.empty() ? Position(-1) : b.m_moveRecord.back(); })
        .def_property_readonly("move_counts", [](const Board& b) {
            py::dict d;
            for (auto p : { Player::Black, Player::None, Player::White }) d[py::cast(p)] = b.moveCounts(p);
            return d;
        })
        .def_property_readonly("move_states", [](const Board& b) {
            py::dict d;
            for (auto p : { Player::Black, Player::None, Player::White }) {
                py::array_t<std::uint8_t> a({ int(HEIGHT), int(WIDTH) });
                for (int c = 0; c < BOARD_SIZE; ++c) a.mutable_data()[c] = b.moveState(p, c);
                d[py::cast(p)] = a;
            }
            return d;
        })
        .def_property_readonly("status", [](const Board& b) {
            auto s = b.status();
            return py::dict("is_end"_a = s.end, "cur_player"_a = s.curPlayer, "winner"_a = s.winner);
        })
        .def("encoded_states", [](const Board& b) {          // game_ext.hpp:87-104: [X_t, Y_t, Z_t, last, last-1, C<is_black>]
            py::array_t<std::uint8_t> st({ 6, int(HEIGHT), int(WIDTH) });
            std::uint8_t* d = st.mutable_data();
            int plane = 0;
            for (auto p : { b.m_curPlayer, -b.m_curPlayer, Player::None }) {
                for (int c = 0; c < BOARD_SIZE; ++c) d[plane * BOARD_SIZE + c] = b.moveState(p, c);
                ++plane;
            }
            for (int i = 0; i <= 1; ++i, ++plane) {
                std::fill(d + plane * BOARD_SIZE, d + (plane + 1) * BOARD_SIZE, std::uint8_t(0));
                if (b.m_moveRecord.size() > std::size_t(i)) d[plane * BOARD_SIZE + (b.m_moveRecord.rbegin() + i)->id] = 1;
            }
            std::fill(d + plane * BOARD_SIZE, d + (plane + 1) * BOARD_SIZE, std::uint8_t(b.m_curPlayer == Player::Black));
            return st;
        }, "Feature planes: [X_t, Y_t, Z_t, y_t-1, x_t-2, C<is_black>]")
        .def("packed", [](const Board& b) {
            py::array_t<std::uint32_t> a(16);
            b.pack(a.mutable_data());
            return a;
        }, "The 64-byte 2-bit image the GPU entry points take.")
        .def("__repr__", [](const Board& b) { return std::string("Board(cur_player: ") + player_str(b.m_curPlayer) + ")"; })
        .def("__str__", &Board::toString);

    // ---- mcts_ext.hpp --------------------------------------------------------------------------------------
    py::class_<Node>(mod, "Node", "MCTS Tree Node")
        .def(py::init<Node*, Position, Player, float, float>(), "parent"_a = nullptr, "position"_a = Position(-1),
             "player"_a = Player::None, "state_value"_a = 0.0, "action_prob"_a = 0.0)
        .def_readonly("parent", &Node::parent)
        .def_readonly("position", &Node::position)
        .def_readonly("player", &Node::player)
        .def_readwrite("state_value", &Node::state_value)
        .def_readwrite("action_prob", &Node::action_prob)
        .def_readwrite("node_visits", &Node::node_visits)
        .def_property_readonly("children", [](const Node* n) {
            py::list children(n->children.size());
            for (std::size_t i = 0; i < n->children.size(); ++i) children[i] = py::cast(n->children[i].get(), py::return_value_policy::reference);
            return children;
        })
        .def("is_leaf", &Node::isLeaf)
        .def("is_full", &Node::isFull)
        .def("__repr__", [](const Node* n) {
            return "Node(pose: " + pos_str(n->position) + ", player: " + player_str(n->player) + ", value: " + std::to_string(n->state_value) +
                   ", prob: " + std::to_string(n->action_prob) + ", visits: " + std::to_string(n->node_visits) + ", childs: " + std::to_string(n->children.size()) + ")";
        });

    py::class_<Policy, std::shared_ptr<Policy>>(mod, "Policy", "MCTS Tree Policy")
        .def(py::init<Policy::SelectFunc, Policy::ExpandFunc, Policy::EvalFunc, Policy::UpdateFunc, double>(), "select"_a = nullptr,
             "expand"_a = nullptr, "eval_state"_a = nullptr, "back_prop"_a = nullptr, "c_puct"_a = C_PUCT)
        .def("prepare", &Policy::prepare)
        .def("clean_up", &Policy::cleanup)
        .def("apply_move", &Policy::applyMove)
        .def("revert_move", &Policy::revertMove)
        .def("check_game_end", &Policy::checkGameEnd)
        .def("create_node", &Policy::createNode)
        .def_readonly("select", &Policy::select)
        .def_readonly("expand", &Policy::expand)
        .def_readonly("eval_state", &Policy::simulate)
        .def_readonly("back_prop", &Policy::backPropogate)
        .def_readonly("c_puct", &Policy::c_puct)
        .def("__repr__", [](const Policy& p) { return "Policy(c_puct: " + std::to_string(p.c_puct) + ", init_acts: " + std::to_string(p.m_initActs) + ")"; });

    py::class_<MCTS>(mod, "MCTS", "Monte Carlo Tree Search")
        .def(py::init<milliseconds, Position, Player, std::shared_ptr<Policy>>(), "c_duration"_a = milliseconds(960),
             "last_move"_a = Position(-1), "last_player"_a = Player::White, py::arg_v("policy", nullptr, "Default Policy"))
        .def(py::init<std::size_t, Position, Player, std::shared_ptr<Policy>>(), "c_iterations"_a, "last_move"_a = Position(-1),
             "last_player"_a = Player::White, py::arg_v("policy", nullptr, "Default Policy"))
        .def_readonly("size", &MCTS::m_size)
        .def_readonly("iterations", &MCTS::m_iterations)
        .def_readonly("duration", &MCTS::m_duration)
        .def_property_readonly("root", [](const MCTS& m) { return m.m_root.get(); }, py::return_value_policy::reference)
        .def_property_readonly("policy", [](const MCTS& m) { return m.m_policy; })
        .def("get_action", &MCTS::getAction)
        .def("eval_state", [](MCTS& m, Board& b) { auto [v, p] = m.evalState(b); return py::make_tuple(v, to_numpy(p)); })
        .def("step_forward", [](MCTS& m) { m.stepForward(); })
        .def("step_forward", [](MCTS& m, Position p) { m.stepForward(p); }, "next_move"_a)
        .def("sync_with_board", &MCTS::syncWithBoard)
        .def("reset", &MCTS::reset)
        .def("__repr__", [](const MCTS& m) { return std::string("MCTS(root_player: ") + player_str(m.m_root->player) + ", nodes: " + std::to_string(m.m_size) + ")"; });

    // ---- policy_ext.hpp -------------------------------------------------------------------------------------
    py::class_<RandomPolicy, Policy, std::shared_ptr<RandomPolicy>>(mod, "RandomPolicy", "Random policy with averaged multiple rollouts (GPU rollout kernel)")
        .def(py::init<double, std::size_t>(), "c_puct"_a = C_PUCT, "c_rollouts"_a = 5)
        .def_readonly("c_rollouts", &RandomPolicy::c_rollouts)
        .def("__repr__", [](const RandomPolicy& p) {
            return "RandomPolicy(c_puct: " + std::to_string(p.c_puct) + ", c_rollouts: " + std::to_string(p.c_rollouts) + ", init_acts: " + std::to_string(p.m_initActs) + ")";
        });

    py::class_<PoolRAVEPolicy, Policy, std::shared_ptr<PoolRAVEPolicy>>(mod, "PoolRAVEPolicy", "PoolRAVE policy with MC-RAVE algorithm (GPU playouts)")
        .def(py::init<double, double>(), "c_puct"_a = 2, "c_bias"_a = 0)                  // policy_ext.hpp:25-36
        .def_readonly("c_bias", &PoolRAVEPolicy::c_bias)
        .def("__repr__", [](const PoolRAVEPolicy& p) {
            return "PoolRAVEPolicy(c_puct: " + std::to_string(p.c_puct) + ", c_bias: " + std::to_string(p.c_bias) + ", init_acts: " + std::to_string(p.m_initActs) + ")";
        });

    py::class_<TraditionalPolicy, Policy, std::shared_ptr<TraditionalPolicy>>(mod, "TraditionalPolicy", "Traditional policy with MC + Pattern Matching algorithm (GPU pattern evaluator)")
        .def(py::init<double, double, bool>(), "c_puct"_a = C_PUCT, "c_bias"_a = 0, "use_rave"_a = false)
        .def("__repr__", [](const TraditionalPolicy& p) {
            return "TraditionalPolicy(c_puct: " + std::to_string(p.c_puct) + (p.c_useRave ? ", c_bias: " + std::to_string(p.c_bias) : std::string()) +
                   ", init_acts: " + std::to_string(p.m_initActs) + ")";
        });

    // ---- new: root-parallel search -------------------------------------------------------------------------------
    py::class_<RootParallelSearch>(mod, "RootParallelSearch", "Root-parallel MCTS: many trees, leaves simulated in one GPU batch per round")
        .def(py::init([](int trees, int c_rollouts, double c_puct, std::uint64_t seed, int replica_base, int threads, bool noise, bool eager, int groups, bool watch) {
                 RootParallelConfig cfg;
                 cfg.trees = trees; cfg.c_rollouts = c_rollouts; cfg.c_puct = c_puct; cfg.seed = seed;
                 cfg.replica_base = replica_base; cfg.threads = threads; cfg.noise = noise; cfg.eager = eager; cfg.groups = groups; cfg.watch = watch;
                 return new RootParallelSearch(cfg);
             }), "trees"_a = 256, "c_rollouts"_a = 5, "c_puct"_a = C_PUCT, "seed"_a = 1, "replica_base"_a = 0, "threads"_a = 0, "noise"_a = false,
             "eager"_a = false, "groups"_a = 0, "watch"_a = true)
        .def("run", [](RootParallelSearch& s, const Board& b, int playouts_per_tree, py::object seed) {
            const bool reseed = !seed.is_none();
            const std::uint64_t key = reseed ? seed.cast<std::uint64_t>() : 0;
            { py::gil_scoped_release release; if (reseed) s.run(b, playouts_per_tree, key); else s.run(b, playouts_per_tree); }
            py::array_t<std::int64_t> a({ 3, int(BOARD_SIZE) });
            std::copy(s.stats().begin(), s.stats().end(), a.mutable_data());
            return a;
        }, "board"_a, "playouts_per_tree"_a, "seed"_a = py::none(), "Returns int64[3,225]: visits, black-won rollouts, white-won rollouts per root move.")
        .def_static("best_move", [](py::array_t<std::int64_t, py::array::c_style | py::array::forcecast> a) {
            if (a.size() != 3 * BOARD_SIZE) throw std::invalid_argument("stats must be int64[3,225]");
            RootParallelSearch::Stats st;
            std::copy(a.data(), a.data() + 3 * BOARD_SIZE, st.begin());
            return RootParallelSearch::bestMove(st);
        })
        .def("tree_dump", [](const RootParallelSearch& s, int tree) {
            const auto nodes = s.dumpTree(tree);
            const py::ssize_t n = static_cast<py::ssize_t>(nodes.size());
            py::array_t<std::int16_t> pos(n), depth(n);
            py::array_t<std::int32_t> visits(n), n_moves(n);
            py::array_t<float> value(n), prior(n);
            for (py::ssize_t i = 0; i < n; ++i) {
                pos.mutable_data()[i] = nodes[i].position; depth.mutable_data()[i] = nodes[i].depth;
                visits.mutable_data()[i] = nodes[i].visits; n_moves.mutable_data()[i] = nodes[i].n_moves;
                value.mutable_data()[i] = nodes[i].value; prior.mutable_data()[i] = nodes[i].prior;
            }
            return py::dict("pos"_a = pos, "visits"_a = visits, "value"_a = value, "prior"_a = prior, "depth"_a = depth, "n_children"_a = n_moves);
        }, "tree"_a, "Visited nodes of one tree of the last run in pre-order (for node-for-node comparison with the reference's MCTS).")
        .def_readonly("seconds_total", &RootParallelSearch::seconds_total)
        .def_readonly("seconds_gpu", &RootParallelSearch::seconds_gpu)
        .def_property_readonly("driver_seconds", [](const RootParallelSearch& s) { return py::make_tuple(s.driver_seconds[0], s.driver_seconds[1], s.driver_seconds[2]); })
        .def_readonly("leaves", &RootParallelSearch::leaves)
        .def_readonly("nodes", &RootParallelSearch::nodes);
}
