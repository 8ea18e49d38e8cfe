// C ABI of the batched self-play search (include/caz_mcts.h) over chess/board.hpp and chess/mcts.hpp.
#include "caz_mcts.h"

#include <cstring>
#include <new>
#include <sstream>
#include <string>

#include "chess/mcts.hpp"

using namespace caz::chess;

struct caz_mcts {
  Mcts* impl;
};

extern "C" {

int caz_mcts_create(const caz_mcts_config* c, int32_t n_games, uint64_t seed, const int32_t* move_label,
                    const int32_t* unflipped_index, int32_t n_threads, caz_mcts** out) {
  if (!c || !out || !move_label || !unflipped_index || n_games < 1 || c->n_labels < 1 || c->search_threads < 1 ||
      c->simulation_num_per_move < 2)   /* with 1 only the root is expanded: no visit counts to pick a move from */
    return 1;
  MctsConfig m;
  m.simulation_num_per_move = c->simulation_num_per_move;
  m.search_threads = c->search_threads;
  m.c_puct = c->c_puct;
  m.noise_eps = c->noise_eps;
  m.dirichlet_alpha = c->dirichlet_alpha;
  m.tau_decay_rate = c->tau_decay_rate;
  m.virtual_loss = c->virtual_loss;
  m.resign_enabled = c->resign_enabled;
  m.resign_threshold = c->resign_threshold;
  m.min_resign_turn = c->min_resign_turn;
  m.max_game_length = c->max_game_length;
  m.n_labels = c->n_labels;
  m.record_play_data = c->record_play_data;
  // no exception may cross the C boundary: allocation or thread-creation failures become status 2
  try {
    caz_mcts* h = new caz_mcts;
    h->impl = nullptr;
    try {
      h->impl = new Mcts(m, n_games, seed, move_label, unflipped_index, n_threads);
    } catch (...) {
      delete h;
      throw;
    }
    h->impl->continuous = c->continuous != 0;
    *out = h;
    return 0;
  } catch (...) {
    return 2;
  }
}

void caz_mcts_destroy(caz_mcts* m) {
  if (!m) return;
  delete m->impl;
  delete m;
}

int caz_mcts_collect(caz_mcts* m, uint8_t* records, int32_t capacity, int32_t* n_leaves) {
  if (!m || !records || !n_leaves) return 1;
  int n = 0;
  int rc;
  try {
    rc = m->impl->collect(records, capacity, &n);
  } catch (...) {
    rc = 2;
  }
  *n_leaves = n;
  return rc;
}

int caz_mcts_collect_csr(caz_mcts* m, uint8_t* records, int32_t capacity, int32_t* n_leaves, int32_t* offsets, int32_t* labels,
                         int32_t edge_capacity) {
  if (!m || !records || !n_leaves || !offsets || !labels) return 1;
  int n = 0;
  int rc;
  try {
    rc = m->impl->collect(records, capacity, &n, offsets, labels, edge_capacity);
  } catch (...) {
    rc = 2;
  }
  *n_leaves = n;
  return rc;
}

int caz_mcts_apply_priors(caz_mcts* m, const double* priors, const float* value, int32_t n) {
  if (!m || (n > 0 && (!priors || !value))) return 1;
  try {
    return m->impl->apply_priors(priors, value, n);
  } catch (...) {
    return 2;
  }
}

int caz_mcts_apply(caz_mcts* m, const float* policy, const float* value, int32_t n) {
  if (!m || (n > 0 && (!policy || !value))) return 1;
  try {
    return m->impl->apply(policy, value, n);
  } catch (...) {
    return 2;
  }
}

int caz_mcts_get_stats(const caz_mcts* m, caz_mcts_stats* out) {
  if (!m || !out) return 1;
  const MctsStats& s = m->impl->stats;
  out->descents = s.descents; out->leaves_evaluated = s.leaves_evaluated; out->terminal_hits = s.terminal_hits;
  out->withdrawn = s.withdrawn; out->moves_played = s.moves_played; out->games_finished = s.games_finished;
  out->white_wins = s.white_wins; out->black_wins = s.black_wins; out->draws = s.draws;
  out->resignations = s.resignations; out->adjudications = s.adjudications; out->tree_nodes_last = s.tree_nodes_last;
  return 0;
}

int caz_mcts_set_position(caz_mcts* m, int32_t game, const char* fen) {
  if (!m || !fen || game < 0 || game >= static_cast<int32_t>(m->impl->games.size())) return 1;
  Position p;
  if (!p.from_fen(fen)) return 1;
  GameSearch& gs = m->impl->games[game];
  gs.game.reset(p);
  gs.finished = false;
  gs.winner = 0;
  gs.begin_move();
  return 0;
}

static int copy_out(const std::string& s, char* buf, int32_t cap) {
  if (!buf || static_cast<int32_t>(s.size()) + 1 > cap) return -1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return static_cast<int>(s.size());
}

int caz_mcts_game_fen(const caz_mcts* m, int32_t game, char* buf, int32_t cap) {
  if (!m || game < 0 || game >= static_cast<int32_t>(m->impl->games.size())) return -1;
  return copy_out(m->impl->games[game].game.pos.fen(), buf, cap);
}

int caz_mcts_game_result(const caz_mcts* m, int32_t game) {
  if (!m || game < 0 || game >= static_cast<int32_t>(m->impl->games.size())) return -1;
  const GameSearch& gs = m->impl->games[game];
  return gs.finished ? gs.winner : 0;
}

int caz_mcts_root_edges(const caz_mcts* m, int32_t game, int32_t* label, int32_t* n, double* q, double* p, int32_t cap) {
  if (!m || game < 0 || game >= static_cast<int32_t>(m->impl->games.size())) return -1;
  const GameSearch& gs = m->impl->games[game];
  const int32_t found = gs.index.find(gs.game.pos.state_key());
  if (found < 0 || gs.nodes[found].pending) return -1;
  const Node& nd = gs.nodes[found];
  int k = 0;
  for (; k < nd.n_edges && k < cap; ++k) {
    const Edge& e = gs.edges[nd.first_edge + k];
    if (label) label[k] = m->impl->move_label[(move_from(e.move) * 64 + move_to(e.move)) * 5 + move_promo(e.move)];
    if (n) n[k] = e.n;
    if (q) q[k] = e.q;
    if (p) p[k] = e.p;
  }
  return k;
}

int64_t caz_mcts_drain_play_data(caz_mcts* m, uint8_t* buf, int64_t cap) {
  if (!m) return -1;
  std::lock_guard<std::mutex> lk(m->impl->play_data_mu);
  const int64_t need = static_cast<int64_t>(m->impl->play_data.size());
  if (need > cap || (need > 0 && !buf)) return need;
  if (need > 0) std::memcpy(buf, m->impl->play_data.data(), static_cast<size_t>(need));
  m->impl->play_data.clear();
  return need;
}

int caz_mcts_debug_dirichlet(uint64_t seed, double alpha, int32_t n, int32_t draws, double* out) {
  if (!out || n < 1 || n > 256 || draws < 1 || alpha <= 0) return 1;
  FastRng rng;
  rng.seed(seed * 0x9E3779B97F4A7C15ull + 1);
  for (int32_t d = 0; d < draws; ++d) {
    double tot = 0.0;
    for (int32_t i = 0; i < n; ++i) { out[(size_t)d * n + i] = rng.gamma(static_cast<float>(alpha)); tot += out[(size_t)d * n + i]; }
    for (int32_t i = 0; i < n; ++i) out[(size_t)d * n + i] /= tot > 0 ? tot : 1.0;
  }
  return 0;
}

uint64_t caz_chess_perft(const char* fen, int32_t depth) {
  Position p;
  if (!fen || !p.from_fen(fen)) return 0;
  return perft(p, depth);
}

int caz_chess_legal_moves(const char* fen, char* buf, int32_t cap) {
  Position p;
  if (!fen || !p.from_fen(fen)) return -1;
  Move m[256];
  const int n = p.legal_moves(m);
  std::string s;
  for (int i = 0; i < n; ++i) {
    if (i) s.push_back(' ');
    s += move_uci(m[i]);
  }
  return copy_out(s, buf, cap) < 0 ? -1 : n;
}

int caz_chess_play(const char* fen, const char* moves, char* fen_out, int32_t cap) {
  Position p;
  if (!fen || !p.from_fen(fen)) return -1;
  Game g;
  g.reset(p);
  std::istringstream in(moves ? moves : "");
  std::string tok;
  while (in >> tok) {
    Move legal[256];
    const int n = g.pos.legal_moves(legal);
    const Move want = move_from_uci(tok);
    bool ok = false;
    for (int i = 0; i < n; ++i) ok |= legal[i] == want;
    if (!ok) return -1;
    g.push(want);
  }
  if (fen_out && copy_out(g.pos.fen(), fen_out, cap) < 0) return -1;
  return g.result();
}

int caz_chess_record(const char* fen, uint8_t* record68) {
  Position p;
  if (!fen || !record68 || !p.from_fen(fen)) return 1;
  p.record(record68);
  return 0;
}

}  // extern "C"
