// game.h -- host-side mirror of the reference's Game layer (include/Game.h, src/Game.cpp):
// Player, Position, Board with the same public surface, so code written against the reference's
// Board / Policy / MCTS plugin API compiles against this one.  Storage is one byte per cell plus
// the packed 2-bit image the GPU entry points take; semantics follow the cited reference lines.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace gomoku {

enum GameConfig { WIDTH = 15, HEIGHT = 15, MAX_RENJU = 5, BOARD_SIZE = WIDTH * HEIGHT };   // Game.h:12-15

enum class Player : short { White = -1, None = 0, Black = 1 };                             // Game.h:19-21
constexpr Player operator-(Player p) { return Player(-static_cast<short>(p)); }
constexpr float CalcScore(Player player, Player winner) { return static_cast<float>(player) * static_cast<float>(winner); }
constexpr float CalcScore(Player player, float value) { return static_cast<float>(player) * value; }

struct Position {                                                                           // Game.h:45-56
    short id;
    constexpr Position(int id = -1) : id(static_cast<short>(id)) {}
    constexpr Position(int x, int y) : id(static_cast<short>(y * WIDTH + x)) {}
    constexpr operator int() const { return id; }
    constexpr int x() const { return id % WIDTH; }
    constexpr int y() const { return id / WIDTH; }
};

class Board {
public:
    Board() { reset(); }

    // Game.cpp:37-47: an invalid move is a no-op that returns the same player
    Player applyMove(Position move, bool checkVictory = true) {
        if (m_curPlayer != Player::None && checkMove(move)) {
            m_cells[move.id] = m_curPlayer == Player::Black ? 1 : 2;
            m_packed[move.id >> 4] |= std::uint32_t(m_cells[move.id]) << ((move.id & 15) * 2);
            m_counts[idx(m_curPlayer)] += 1;
            m_counts[idx(Player::None)] -= 1;
            m_moveRecord.push_back(move);
            m_curPlayer = -m_curPlayer;
            if (checkVictory) checkGameEnd();
        }
        return m_curPlayer;
    }

    Player revertMove(std::size_t count = 1) {                                              // Game.cpp:49-62
        if (m_curPlayer == Player::None && count != 0) {
            m_curPlayer = moveCounts(Player::Black) == moveCounts(Player::White) ? Player::Black : Player::White;
            m_winner = Player::None;
        }
        for (std::size_t i = 0; !m_moveRecord.empty() && i < count; ++i) {
            m_cells[m_moveRecord.back().id] = 0;
            m_packed[m_moveRecord.back().id >> 4] &= ~(std::uint32_t(3) << ((m_moveRecord.back().id & 15) * 2));
            m_counts[idx(-m_curPlayer)] -= 1;
            m_counts[idx(Player::None)] += 1;
            m_moveRecord.pop_back();
            m_curPlayer = -m_curPlayer;
        }
        return m_curPlayer;
    }

    // Game.cpp:64-73: uniform start index, then the first empty cell at or after it, cyclically
    Position getRandomMove() const {
        if (moveCounts(Player::None) == 0) throw std::overflow_error("board is already full");
        static thread_local std::mt19937 engine{ std::random_device{}() };
        int id = static_cast<int>(std::uniform_int_distribution<unsigned>(0, BOARD_SIZE - 1)(engine));
        while (m_cells[id] != 0) id = (id + 1) % BOARD_SIZE;
        return Position(id);
    }

    bool checkMove(Position move) const { return move.id >= 0 && move.id < BOARD_SIZE && m_cells[move.id] == 0; }
    static bool checkBoundary(int x, int y) { return x >= 0 && x < WIDTH && y >= 0 && y < HEIGHT; }

    bool checkGameEnd() {                                                                   // Game.cpp:88-136
        if (m_curPlayer == Player::None) return true;
        if (m_moveRecord.empty()) return false;
        const int cx = m_moveRecord.back().x(), cy = m_moveRecord.back().y();
        const Player last = -m_curPlayer;
        const std::uint8_t stone = last == Player::Black ? 1 : 2;
        static constexpr int DX[4] = { 1, 0, 1, 1 }, DY[4] = { 0, 1, -1, 1 };
        for (int d = 0; d < 4; ++d) {
            int run = 1;
            for (int sgn = 1; sgn >= -1; sgn -= 2)
                for (int i = 1, x = cx + sgn * DX[d], y = cy + sgn * DY[d]; i <= MAX_RENJU; ++i, x += sgn * DX[d], y += sgn * DY[d]) {
                    if (!checkBoundary(x, y) || m_cells[y * WIDTH + x] != stone) break;
                    ++run;
                }
            if (run >= MAX_RENJU) {
                m_winner = last;
                m_curPlayer = Player::None;
                return true;
            }
        }
        if (moveCounts(Player::None) == 0) {
            m_winner = Player::None;
            m_curPlayer = Player::None;
            return true;
        }
        return false;
    }

    void reset() {                                                                          // Game.cpp:138-146
        m_cells.fill(0);
        m_packed.fill(0);
        m_counts = { 0, BOARD_SIZE, 0 };
        m_moveRecord.clear();
        m_curPlayer = Player::Black;
        m_winner = Player::None;
    }

    struct Status { bool end; Player curPlayer; Player winner; };
    Status status() const { return { m_curPlayer == Player::None, m_curPlayer, m_winner }; }   // Game.h:100-105
    bool moveState(Player player, Position pose) const {                                    // Game.h:112
        return m_cells[pose.id] == (player == Player::Black ? 1 : player == Player::White ? 2 : 0);
    }
    std::size_t moveCounts(Player player) const { return m_counts[idx(player)]; }            // Game.h:115-116
    std::uint8_t cell(int id) const { return m_cells[id]; }                                  // 0 empty, 1 black, 2 white

    // the 64-byte image the GPU entry points take (include/gomoku_b200.h), kept up to date by applyMove / revertMove
    void pack(std::uint32_t out[16]) const {
        for (int i = 0; i < 16; ++i) out[i] = m_packed[i];
    }

    std::string toString() const {                                                          // Game.cpp:177-205
        std::string s = "  ";
        static const char* hex = "0123456789abcde";
        for (int x = 0; x < WIDTH; ++x) { s += hex[x]; s += ' '; }
        s += '\n';
        for (int y = 0; y < HEIGHT; ++y) {
            s += hex[y]; s += ' ';
            for (int x = 0; x < WIDTH; ++x) { s += m_cells[y * WIDTH + x] == 1 ? 'x' : m_cells[y * WIDTH + x] == 2 ? 'o' : '_'; s += ' '; }
            s += '\n';
        }
        return s;
    }

    bool operator==(const Board& o) const {
        return m_cells == o.m_cells && m_curPlayer == o.m_curPlayer && m_winner == o.m_winner && m_moveRecord.size() == o.m_moveRecord.size();
    }

public:
    Player m_curPlayer = Player::Black;          // Player::None once the game has ended (Game.h:123-128)
    Player m_winner = Player::None;
    std::vector<Position> m_moveRecord;

private:
    static int idx(Player p) { return static_cast<int>(p) + 1; }
    std::array<std::uint8_t, BOARD_SIZE> m_cells{};
    std::array<std::uint32_t, 16> m_packed{};    // 2 bits per cell, cell c in word c / 16 at bits 2 (c % 16)
    std::array<std::size_t, 3> m_counts{};
};

}  // namespace gomoku
