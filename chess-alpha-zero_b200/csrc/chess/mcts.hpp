// Batched AlphaZero search for many concurrent self-play games, host side.
//
// What it restates: ChessPlayer.action / search_moves / search_my_move / select_action_q_and_u / calc_policy /
// apply_temperature (reference agent/player_chess.py:119-333) and the game loop of self_play_buffer
// (worker/self_play.py:113-154) with ChessEnv.step / _game_over / _resign (env/chess_env.py:79-118).
//
// What is different, on purpose: the reference runs `search_threads` Python threads per game, each blocking on a pipe
// for its leaf's evaluation.  Here one call (`collect`) walks up to `search_threads` descents per game, one after the
// other with the reference's virtual loss applied between them, for ALL games, and returns the leaves as 68-byte board
// records; the caller evaluates them in ONE batch (caz_eval_records) and hands policy/value back (`apply`), which
// expands the leaves, backs the values up and -- once a game has its `simulation_num_per_move` descents -- plays the
// move.  Same arithmetic per descent (PUCT score, virtual loss, backup, prior normalisation, temperature), same tree
// semantics (a fresh tree per move keyed by the FEN without the move number, terminal test on the path's own history);
// the interleaving of a game's descents is sequential instead of racy.  A descent that runs into a leaf another
// descent of the same batch is still waiting for (the reference thread would block on node_lock there) is withdrawn
// and the game's batch ends early.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <string>
#include <vector>

#include "board.hpp"

namespace caz {
namespace chess {

struct MctsConfig {
  int simulation_num_per_move = 800;   // configs/normal.py:33
  int search_threads = 16;             // :31
  double c_puct = 1.5;                 // :36
  double noise_eps = 0.25;             // :37
  double dirichlet_alpha = 0.3;        // :38
  double tau_decay_rate = 0.99;        // :39
  int virtual_loss = 3;                // :40
  double resign_threshold = -0.8;      // :41
  int resign_enabled = 1;
  int min_resign_turn = 5;             // :42
  int max_game_length = 1000;          // :43
  int n_labels = 1968;
  int record_play_data = 0;            // keep [fen, policy, z] rows of finished games (self_play.py:90-100, 135-152)
};

struct MctsStats {
  int64_t descents, leaves_evaluated, terminal_hits, withdrawn, moves_played, games_finished;
  int64_t white_wins, black_wins, draws, resignations, adjudications, tree_nodes_last;
};

constexpr int MAX_MOVES = 256;   // bound on the legal moves of a position (218 is the known maximum)

struct Edge {
  Move move;
  int32_t label;   // index into the policy vector the network returned for this node (flip already applied)
  int32_t n;
  double w, q, p;
};
struct Node {
  int32_t first_edge, n_edges;
  int32_t sum_n;
  bool pending;   // evaluation requested, priors not there yet
};

struct Request {
  int32_t node;
  int32_t path_begin, path_len;   // into Search::path_pool: (node, edge) pairs from the root down
};

// Per-game random numbers: xoshiro256++ with single-precision Gamma draws for the root's Dirichlet noise.  The reference
// draws np.random.dirichlet([alpha] * n) afresh in EVERY root selection (player_chess.py:285-286), 800 times per move:
// with std::gamma_distribution that was a third of the search's time.
struct FastRng {
  uint64_t s[4];
  float spare = 0.0f;
  bool has_spare = false;
  void seed(uint64_t x) {
    for (auto& w : s) {   // splitmix64
      uint64_t z = (x += 0x9E3779B97F4A7C15ull);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      w = z ^ (z >> 31);
    }
    has_spare = false;
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0, 1)
  float uniform_open() { return (static_cast<float>(next() >> 40) + 0.5f) * (1.0f / 16777216.0f); }  // (0, 1)
  float normal() {   // Marsaglia's polar method, values in pairs
    if (has_spare) { has_spare = false; return spare; }
    for (;;) {
      const float u = 2.0f * uniform_open() - 1.0f, v = 2.0f * uniform_open() - 1.0f, q = u * u + v * v;
      if (q >= 1.0f || q == 0.0f) continue;
      const float f = std::sqrt(-2.0f * std::log(q) / q);
      spare = v * f;
      has_spare = true;
      return u * f;
    }
  }
  // Gamma(alpha, 1): Marsaglia & Tsang (2000); for alpha < 1 the usual Gamma(alpha + 1) * U^(1/alpha)
  float gamma(float alpha) {
    const float a = alpha < 1.0f ? alpha + 1.0f : alpha;
    const float d = a - 1.0f / 3.0f, c = 1.0f / std::sqrt(9.0f * d);
    float g;
    for (;;) {
      const float x = normal(), t = 1.0f + c * x;
      if (t <= 0.0f) continue;
      const float v = t * t * t, u = uniform_open(), x2 = x * x;
      // the paper's squeeze accepts ~92 % of the candidates without the two logarithms
      if (u < 1.0f - 0.0331f * x2 * x2 || std::log(u) < 0.5f * x2 + d - d * v + d * std::log(v)) { g = d * v; break; }
    }
    if (alpha < 1.0f) g *= std::exp(std::log(uniform_open()) / alpha);
    return g;
  }
};

// state key -> node index, open addressing (a tree holds at most simulation_num_per_move + 1 nodes)
class NodeIndex {
 public:
  void clear(size_t expected) {
    size_t cap = 64;
    while (cap < expected * 2 + 2) cap <<= 1;
    keys_.assign(cap, 0);
    vals_.assign(cap, -1);
    used_ = 0;
  }
  int32_t find(uint64_t key) const {
    const size_t mask = keys_.size() - 1;
    for (size_t i = key & mask;; i = (i + 1) & mask) {
      if (vals_[i] < 0) return -1;
      if (keys_[i] == key) return vals_[i];
    }
  }
  void emplace(uint64_t key, int32_t val) {
    if ((used_ + 1) * 2 > keys_.size()) grow();
    insert(key, val);
  }

 private:
  void insert(uint64_t key, int32_t val) {
    const size_t mask = keys_.size() - 1;
    size_t i = key & mask;
    while (vals_[i] >= 0) i = (i + 1) & mask;
    keys_[i] = key;
    vals_[i] = val;
    ++used_;
  }
  void grow() {
    std::vector<uint64_t> k;
    std::vector<int32_t> v;
    k.swap(keys_);
    v.swap(vals_);
    keys_.assign(k.size() * 2, 0);
    vals_.assign(k.size() * 2, -1);
    used_ = 0;
    for (size_t i = 0; i < k.size(); ++i)
      if (v[i] >= 0) insert(k[i], v[i]);
  }
  std::vector<uint64_t> keys_;
  std::vector<int32_t> vals_;
  size_t used_ = 0;
};

// A fixed pool of worker threads that runs f(chunk) for chunk = 0..n_chunks-1 (one call at a time).
class WorkerPool {
 public:
  explicit WorkerPool(int n) {
    for (int i = 0; i < n; ++i) threads_.emplace_back([this, i]() { loop(i); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int size() const { return static_cast<int>(threads_.size()); }
  void run(const std::function<void(int)>& f) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &f;
      pending_ = size();
      ++generation_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this]() { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* f;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        f = fn_;
      }
      (*f)(id);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

class GameSearch {
 public:
  Game game;
  FastRng rng;
  // the tree of the move being searched (ChessPlayer.reset() clears it for every move, player_chess.py:131)
  NodeIndex index;
  std::vector<Node> nodes;
  std::vector<Edge> edges;
  std::vector<std::pair<int32_t, int32_t>> path_pool;
  std::vector<Request> requests;        // leaves handed out by the last collect()
  std::vector<uint8_t> request_records;
  int sims_done = 0, inflight = 0;
  double root_value_max = -1e30;
  bool finished = false;
  int winner = 0;                       // 1 white, 2 black, 3 draw
  int64_t games_played = 0;
  // ChessPlayer.moves of both players, in game order: observation, visit-count policy (sparse), who moved
  struct MoveRecord { std::string fen; std::vector<std::pair<uint16_t, float>> policy; int mover; };
  std::vector<MoveRecord> move_log;

  void new_game() {
    Position p;
    p.set_startpos();
    game.reset(p);
    finished = false;
    winner = 0;
    move_log.clear();
    begin_move();
  }
  int expected_nodes = 1024;
  void begin_move() {
    index.clear(static_cast<size_t>(expected_nodes));
    nodes.clear();
    edges.clear();
    path_pool.clear();
    requests.clear();
    request_records.clear();
    sims_done = inflight = 0;
    root_value_max = -1e30;
  }
};

struct alignas(128) PaddedStats {
  MctsStats s{};
};

class Mcts {
 public:
  std::unique_ptr<WorkerPool> pool;
  MctsConfig cfg;
  std::vector<GameSearch> games;
  std::vector<int32_t> move_label;        // [64*64*5] -> label or -1 (promotion index: 0 none, 1 n, 2 b, 3 r, 4 q)
  std::vector<int32_t> unflipped_index;   // Config.unflipped_index (config.py:185)
  int n_threads = 1;
  MctsStats stats{};
  std::vector<std::pair<int32_t, int32_t>> batch_map;   // leaf i of the last collect -> (game, request)
  std::vector<int32_t> batch_edge_offsets;               // CSR offsets of the leaves' edges, same order

  Mcts(const MctsConfig& c, int n_games, uint64_t seed, const int32_t* labels, const int32_t* unflip, int threads)
      : cfg(c), games(n_games), move_label(labels, labels + 64 * 64 * 5), unflipped_index(unflip, unflip + c.n_labels),
        n_threads(threads < 1 ? 1 : threads) {
    if (n_threads > n_games) n_threads = n_games;
    if (n_threads > 1) pool.reset(new WorkerPool(n_threads));
    for (int g = 0; g < n_games; ++g) {
      games[g].rng.seed(seed * 0x9E3779B97F4A7C15ull + static_cast<uint64_t>(g) + 1);
      games[g].expected_nodes = c.simulation_num_per_move + 2;
      games[g].new_game();
    }
  }

  int label_of(const Position& pos, Move m) const {
    const int l = move_label[(move_from(m) * 64 + move_to(m)) * 5 + move_promo(m)];
    if (l < 0) return -1;
    // the network sees the position from the side to move; for black the reference permutes the returned policy with
    // Config.flip_policy (player_chess.py:232-233): real[i] = canonical[unflipped_index[i]]
    return pos.turn == WHITE ? l : unflipped_index[l];
  }

  // f(game, stats of the calling worker) for every game: contiguous blocks of games per worker (neighbouring games
  // share cache lines), per-worker statistics merged afterwards
  template <typename F>
  void parallel_games(F&& f) {
    const int n = static_cast<int>(games.size());
    const int t = pool ? pool->size() : 1;
    std::vector<PaddedStats> st(static_cast<size_t>(t));
    if (t <= 1) {
      for (int g = 0; g < n; ++g) f(g, st[0].s);
    } else {
      const std::function<void(int)> job = [&](int w) {
        const int lo = static_cast<int>(static_cast<int64_t>(n) * w / t), hi = static_cast<int>(static_cast<int64_t>(n) * (w + 1) / t);
        for (int g = lo; g < hi; ++g) f(g, st[w].s);
      };
      pool->run(job);
    }
    for (const auto& ps : st) merge(ps.s);
  }

  // select_action_q_and_u (player_chess.py:277-299) on an expanded node; returns the edge index within the node
  int select(GameSearch& gs, const Node& nd, bool is_root) {
    const double xx = std::sqrt(static_cast<double>(nd.sum_n + 1));
    double best_s = -999.0;
    int best = -1;
    double noise[256];
    const bool noisy = is_root && cfg.noise_eps != 0.0;   // with eps = 0 the draw cannot change the choice: skipped
    if (noisy && nd.n_edges > 0) {
      // np.random.dirichlet([alpha] * n): independent Gamma(alpha, 1) draws, normalised -- a fresh draw per call (:285-286)
      const float alpha = static_cast<float>(cfg.dirichlet_alpha);
      double tot = 0.0;
      for (int i = 0; i < nd.n_edges; ++i) { noise[i] = gs.rng.gamma(alpha); tot += noise[i]; }
      if (tot <= 0.0) tot = 1.0;
      for (int i = 0; i < nd.n_edges; ++i) noise[i] /= tot;
    }
    for (int i = 0; i < nd.n_edges; ++i) {
      const Edge& e = gs.edges[nd.first_edge + i];
      double p = e.p;
      if (noisy) p = (1 - cfg.noise_eps) * p + cfg.noise_eps * noise[i];
      const double b = e.q + cfg.c_puct * p * xx / (1 + e.n);
      if (b > best_s) { best_s = b; best = i; }
    }
    return best;
  }

  void backup(GameSearch& gs, int path_begin, int path_len, double leaf_v, MctsStats& st) {
    // search_my_move unwinding (player_chess.py:205-217): negate at every level, undo the virtual loss
    const int vl = cfg.virtual_loss;
    double v = leaf_v;
    for (int i = path_len - 1; i >= 0; --i) {
      v = -v;
      const auto& pe = gs.path_pool[path_begin + i];
      Node& nd = gs.nodes[pe.first];
      Edge& e = gs.edges[nd.first_edge + pe.second];
      nd.sum_n += -vl + 1;
      e.n += -vl + 1;
      e.w += vl + v;
      e.q = e.w / e.n;
    }
    // what the root-level call returned; root_value = np.max(vals) (player_chess.py:158)
    gs.root_value_max = std::max(gs.root_value_max, path_len > 0 ? v : leaf_v);
    ++gs.sims_done;
    ++st.descents;
  }

  void withdraw(GameSearch& gs, int path_begin, int path_len) {
    const int vl = cfg.virtual_loss;
    for (int i = path_len - 1; i >= 0; --i) {
      const auto& pe = gs.path_pool[path_begin + i];
      Node& nd = gs.nodes[pe.first];
      Edge& e = gs.edges[nd.first_edge + pe.second];
      nd.sum_n -= vl;
      e.n -= vl;
      e.w += vl;
      e.q = e.n ? e.w / e.n : 0.0;
    }
    gs.path_pool.resize(path_begin);
  }

  // one game's share of a collect(): descents until search_threads are in flight, the move's quota is reached, or a
  // descent has to be withdrawn
  void collect_game(GameSearch& gs, MctsStats& st) {
    gs.requests.clear();
    gs.request_records.clear();
    gs.path_pool.clear();
    if (gs.finished) return;
    const Position root_pos = gs.game.pos;
    const size_t root_hist = gs.game.hist.size();
    const int root_halfmoves = gs.game.num_halfmoves;
    Move legal[256];
    while (gs.sims_done + gs.inflight < cfg.simulation_num_per_move && gs.inflight < cfg.search_threads) {
      const int path_begin = static_cast<int>(gs.path_pool.size());
      int path_len = 0;
      bool stop_batch = false;
      for (bool is_root = true;; is_root = false) {
        const uint64_t key = gs.game.state_key();
        const int32_t found = gs.index.find(key);
        int n_legal;
        const Move* lm = legal;
        if (found >= 0) n_legal = gs.nodes[found].n_edges;   // a known node's moves are its edges', read only if a claim needs them
        else n_legal = gs.game.pos.legal_moves(legal);
        if (!is_root) {
          // `if env.done` (player_chess.py:176-180): env.step ended the game with result(claim_draw=True)
          const Edge* known = found >= 0 ? gs.edges.data() + gs.nodes[found].first_edge : nullptr;
          const int res = gs.game.result_with([known, lm](int i) { return known ? known[i].move : lm[i]; }, n_legal);
          if (res) {
            ++st.terminal_hits;
            backup(gs, path_begin, path_len, res == 3 ? 0.0 : -1.0, st);
            gs.path_pool.resize(path_begin);
            break;
          }
        }
        if (found < 0) {
          // new leaf: expand_and_evaluate (player_chess.py:186-188, 219-235) -- request its evaluation
          const int32_t id = static_cast<int32_t>(gs.nodes.size());
          Node nd{static_cast<int32_t>(gs.edges.size()), n_legal, 0, true};
          for (int i = 0; i < n_legal; ++i) gs.edges.push_back(Edge{lm[i], label_of(gs.game.pos, lm[i]), 0, 0.0, 0.0, 0.0});
          gs.nodes.push_back(nd);
          gs.index.emplace(key, id);
          gs.requests.push_back(Request{id, path_begin, path_len});
          gs.request_records.resize(gs.request_records.size() + 68);
          gs.game.pos.record(gs.request_records.data() + gs.request_records.size() - 68);
          ++gs.inflight;
          break;
        }
        Node& nd = gs.nodes[found];
        if (nd.pending) {   // the reference thread would wait on node_lock for the evaluation: withdraw, end this batch
          withdraw(gs, path_begin, path_len);
          ++st.withdrawn;
          stop_batch = true;
          break;
        }
        const int a = select(gs, nd, is_root);
        if (a < 0) {   // no legal move at the root (cannot happen in a running game); treat as terminal draw
          backup(gs, path_begin, path_len, 0.0, st);
          gs.path_pool.resize(path_begin);
          break;
        }
        Edge& e = gs.edges[nd.first_edge + a];
        const int vl = cfg.virtual_loss;   // player_chess.py:194-202
        nd.sum_n += vl;
        e.n += vl;
        e.w += -vl;
        e.q = e.w / e.n;
        gs.path_pool.emplace_back(found, a);
        ++path_len;
        gs.game.push(e.move);
      }
      gs.game.pop_to(root_hist, root_pos, root_halfmoves);
      if (stop_batch) break;
    }
  }

  // Runs the descents of every game; writes the leaves' board records (68 bytes each) in game order.
  // Call order is collect -> apply -> collect ...: a second collect() while leaves of the first are still waiting for their
  // apply() returns 1 and changes nothing (collect_game would forget the requests but not the pending nodes, and every
  // later descent would be withdrawn for ever).
  // With `offsets` / `labels` (CSR: offsets[n + 1], labels[offsets[n]]) the leaves' legal-move labels are written too, in
  // edge order: what caz_submit_records_priors needs to send back one prior per legal move instead of the whole policy.
  int collect(uint8_t* records, int capacity, int* n_out, int32_t* offsets = nullptr, int32_t* labels = nullptr,
              int edge_capacity = 0) {
    if (static_cast<int64_t>(games.size()) * cfg.search_threads > capacity) return 1;
    if (offsets && (!labels || static_cast<int64_t>(games.size()) * cfg.search_threads * MAX_MOVES > edge_capacity)) return 1;
    if (!batch_map.empty()) return 1;
    for (const auto& gs : games)
      if (gs.inflight > 0) return 1;
    for (;;) {
      // progress = descents completed + plies played + games ended; an iteration that moves none of them cannot end
      auto progress = [&] {
        int64_t p = 0;
        for (const auto& gs : games) p += gs.sims_done + static_cast<int64_t>(gs.game.num_halfmoves) * 4096 + (static_cast<int64_t>(gs.games_played) << 32) + (gs.finished ? 1 : 0);
        return p;
      };
      const int64_t before = progress();
      parallel_games([&](int g, MctsStats& st) { collect_game(games[g], st); });
      batch_map.clear();
      int n = 0;
      for (size_t g = 0; g < games.size(); ++g) {
        GameSearch& gs = games[g];
        for (size_t r = 0; r < gs.requests.size(); ++r) {
          std::memcpy(records + static_cast<size_t>(n) * 68, gs.request_records.data() + r * 68, 68);
          batch_map.emplace_back(static_cast<int32_t>(g), static_cast<int32_t>(r));
          ++n;
        }
      }
      *n_out = n;
      if (n > 0) {
        batch_edge_offsets.assign(1, 0);
        for (const auto& bm : batch_map) {
          const GameSearch& gs = games[bm.first];
          batch_edge_offsets.push_back(batch_edge_offsets.back() + gs.nodes[gs.requests[bm.second].node].n_edges);
        }
        if (offsets) {
          std::memcpy(offsets, batch_edge_offsets.data(), batch_edge_offsets.size() * sizeof(int32_t));
          int k = 0;
          for (const auto& bm : batch_map) {
            const GameSearch& gs = games[bm.first];
            const Node& nd = gs.nodes[gs.requests[bm.second].node];
            for (int e = 0; e < nd.n_edges; ++e) labels[k++] = gs.edges[nd.first_edge + e].label;
          }
        }
        return 0;
      }
      // nothing to evaluate (every descent ended in a terminal position): finish what can be finished and go on
      bool any_active = false;
      for (auto& gs : games) any_active |= !gs.finished;
      if (!any_active) return 0;
      finish_moves();
      if (progress() == before) return 1;   // cannot happen in a consistent tree; never spin
    }
  }

  // policy: [n][n_labels] network output for the leaves of the last collect (canonical, side to move = white);
  // value: [n].  Expands, backs up, plays the moves of the games whose search is complete.
  int apply(const float* policy, const float* value, int n) {
    return apply_impl(n, value, [&](int i, GameSearch& gs, Node& nd) {
      const float* pol = policy + static_cast<size_t>(i) * cfg.n_labels;
      // push p to edges (player_chess.py:267-275): p = P[label] / (1e-8 + sum of the legal moves' P, in move order)
      double tot = 1e-8;
      for (int e = 0; e < nd.n_edges; ++e) {
        Edge& ed = gs.edges[nd.first_edge + e];
        const float pe = ed.label >= 0 ? pol[ed.label] : 0.0f;
        ed.p = static_cast<double>(pe);
        tot += static_cast<double>(pe);
      }
      for (int e = 0; e < nd.n_edges; ++e) gs.edges[nd.first_edge + e].p /= tot;
    });
  }

  // The same with the prior push already done on the device (caz_submit_records_priors / push_priors_kernel: the same
  // float32 -> float64 arithmetic in the same order, bit-identical): priors[offsets[i] + e] belongs to edge e of leaf i.
  int apply_priors(const double* priors, const float* value, int n) {
    if (static_cast<int>(batch_edge_offsets.size()) != n + 1) return 1;
    return apply_impl(n, value, [&](int i, GameSearch& gs, Node& nd) {
      const double* pr = priors + batch_edge_offsets[i];
      for (int e = 0; e < nd.n_edges; ++e) gs.edges[nd.first_edge + e].p = pr[e];
    });
  }

  template <class PushPriors>
  int apply_impl(int n, const float* value, PushPriors&& push) {
    if (n != static_cast<int>(batch_map.size())) return 1;
    std::vector<int> first(games.size() + 1, 0);
    for (auto& bm : batch_map) ++first[bm.first + 1];
    for (size_t g = 0; g < games.size(); ++g) first[g + 1] += first[g];
    parallel_games([&](int g, MctsStats& st) {
      GameSearch& gs = games[g];
      for (size_t r = 0; r < gs.requests.size(); ++r) {
        const int i = first[g] + static_cast<int>(r);
        const Request& rq = gs.requests[r];
        Node& nd = gs.nodes[rq.node];
        push(i, gs, nd);
        nd.pending = false;
        --gs.inflight;
        ++st.leaves_evaluated;
        backup(gs, rq.path_begin, rq.path_len, static_cast<double>(value[i]), st);
      }
      gs.requests.clear();
    });
    batch_map.clear();
    finish_moves();
    return 0;
  }

  // ChessPlayer.action after search_moves (player_chess.py:131-146) + self_play_buffer's loop body (self_play.py:126-133)
  void finish_moves() {
    parallel_games([&](int g, MctsStats& st) {
      GameSearch& gs = games[g];
      if (gs.finished || gs.inflight > 0 || gs.sims_done < cfg.simulation_num_per_move) return;
      st.tree_nodes_last = static_cast<int64_t>(gs.nodes.size());
      const uint64_t key = gs.game.pos.state_key();
      const Node& root = gs.nodes[gs.index.find(key)];
      // calc_policy (:320-331): visit counts over the label space; apply_temperature (:301-318)
      std::vector<double> pol(cfg.n_labels, 0.0);
      std::vector<Move> move_of(cfg.n_labels, 0);
      double total = 0.0;
      for (int i = 0; i < root.n_edges; ++i) {
        const Edge& e = gs.edges[root.first_edge + i];
        const int l = move_label[(move_from(e.move) * 64 + move_to(e.move)) * 5 + move_promo(e.move)];
        if (l < 0) continue;
        pol[l] = e.n;
        move_of[l] = e.move;
        total += e.n;
      }
      if (total <= 0.0) total = 1.0;   // unreachable: caz_mcts_create rejects fewer than 2 simulations per move
      for (auto& x : pol) x /= total;
      double tau = std::pow(cfg.tau_decay_rate, gs.game.num_halfmoves + 1);
      if (tau < 0.1) tau = 0;
      int action = 0;
      if (tau == 0) {
        for (int l = 1; l < cfg.n_labels; ++l)
          if (pol[l] > pol[action]) action = l;   // np.argmax: first maximum
      } else {
        double sum = 0.0;
        for (auto& x : pol) { x = std::pow(x, 1 / tau); sum += x; }
        const double u = gs.rng.uniform() * sum;
        double acc = 0.0;
        action = -1;
        int last_nonzero = 0;
        for (int l = 0; l < cfg.n_labels; ++l) {
          if (pol[l] <= 0) continue;
          last_nonzero = l;
          acc += pol[l];
          if (u < acc) { action = l; break; }
        }
        if (action < 0) action = last_nonzero;
      }
      const bool resign = cfg.resign_enabled && gs.root_value_max <= cfg.resign_threshold &&
                          gs.game.num_halfmoves > cfg.min_resign_turn;
      if (resign) {
        gs.winner = gs.game.pos.turn == WHITE ? 2 : 1;   // _resign (chess_env.py:108-116)
        ++st.resignations;
        end_game(gs, st);
        return;
      }
      if (cfg.record_play_data) {   // self.moves.append([env.observation, list(policy)]) (player_chess.py:145) -- not on a resignation
        GameSearch::MoveRecord rec;
        rec.fen = gs.game.pos.fen();
        rec.mover = gs.game.pos.turn;
        for (int i = 0; i < root.n_edges; ++i) {
          const Edge& e = gs.edges[root.first_edge + i];
          const int l = move_label[(move_from(e.move) * 64 + move_to(e.move)) * 5 + move_promo(e.move)];
          if (l >= 0 && e.n > 0) rec.policy.emplace_back(static_cast<uint16_t>(l), static_cast<float>(e.n / total));
        }
        gs.move_log.push_back(std::move(rec));
      }
      gs.game.push(move_of[action]);
      ++st.moves_played;
      const int res = gs.game.result();
      // self_play_buffer (self_play.py:127-133): after EVERY step, `if env.num_halfmoves >= max_game_length:
      // env.adjudicate()` -- the material count overrides even a result the step itself produced
      if (gs.game.num_halfmoves >= cfg.max_game_length) {   // adjudicate (chess_env.py:118-129)
        gs.winner = adjudicate(gs.game.pos);
        ++st.adjudications;
        end_game(gs, st);
        return;
      }
      if (res) {
        gs.winner = res;
        end_game(gs, st);
        return;
      }
      gs.begin_move();
    });
  }

  // ChessEnv.adjudicate / testeval(absolute=True) (chess_env.py:118-129,170-188): material balance, white's view
  static int adjudicate(const Position& p) {
    static const double val[6] = {1, 3, 3.25, 5, 14, 3};   // P N B R Q K (piece_vals, chess_env.py:171)
    double ans = 0.0, tot = 0.0;
    for (int pc = 0; pc < 6; ++pc) {
      const double w = popcount(p.by_piece[pc] & p.by_color[WHITE]), b = popcount(p.by_piece[pc] & p.by_color[BLACK]);
      ans += val[pc] * (w - b);
      tot += val[pc] * (w + b);
    }
    const double v = std::tanh(ans / tot * 3);
    if (std::fabs(v) < 0.01) return 3;
    return v > 0 ? 1 : 2;
  }

  // Finished games' rows, serialised: per row  fen '\0' | z int8 | n uint16 | n x (label uint16, probability float32);
  // z as finish_game assigns it (self_play.py:135-143): +1 the mover won, -1 lost, 0 draw.
  std::vector<uint8_t> play_data;
  std::mutex play_data_mu;
  void flush_play_data(GameSearch& gs) {
    std::vector<uint8_t> buf;
    for (const auto& r : gs.move_log) {
      buf.insert(buf.end(), r.fen.begin(), r.fen.end());
      buf.push_back(0);
      const int z = gs.winner == 3 ? 0 : ((gs.winner == 1) == (r.mover == WHITE) ? 1 : -1);
      buf.push_back(static_cast<uint8_t>(static_cast<int8_t>(z)));
      const uint16_t n = static_cast<uint16_t>(r.policy.size());
      const uint8_t* pn = reinterpret_cast<const uint8_t*>(&n);
      buf.insert(buf.end(), pn, pn + 2);
      for (const auto& lp : r.policy) {
        const uint8_t* a = reinterpret_cast<const uint8_t*>(&lp.first);
        const uint8_t* b = reinterpret_cast<const uint8_t*>(&lp.second);
        buf.insert(buf.end(), a, a + 2);
        buf.insert(buf.end(), b, b + 4);
      }
    }
    std::lock_guard<std::mutex> lk(play_data_mu);
    play_data.insert(play_data.end(), buf.begin(), buf.end());
  }

  void end_game(GameSearch& gs, MctsStats& st) {
    if (cfg.record_play_data) flush_play_data(gs);
    ++st.games_finished;
    if (gs.winner == 1) ++st.white_wins;
    else if (gs.winner == 2) ++st.black_wins;
    else ++st.draws;
    ++gs.games_played;
    if (continuous) gs.new_game();
    else gs.finished = true;
  }

  void merge(const MctsStats& s) {
    stats.descents += s.descents; stats.leaves_evaluated += s.leaves_evaluated; stats.terminal_hits += s.terminal_hits;
    stats.withdrawn += s.withdrawn; stats.moves_played += s.moves_played; stats.games_finished += s.games_finished;
    stats.white_wins += s.white_wins; stats.black_wins += s.black_wins; stats.draws += s.draws;
    stats.resignations += s.resignations; stats.adjudications += s.adjudications;
    if (s.tree_nodes_last) stats.tree_nodes_last = s.tree_nodes_last;
  }

  bool continuous = true;   // start a new game when one ends (self-play); false: games stay finished
};

}  // namespace chess
}  // namespace caz
