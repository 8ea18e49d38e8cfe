/*
 * caz_mcts.h -- C ABI of the host-side batched self-play search that feeds the B200 evaluator.
 *
 * Replaces, for many concurrent games at once, the reference's ChessPlayer.action / search_moves / search_my_move /
 * select_action_q_and_u (src/chess_zero/agent/player_chess.py:119-333), the game loop of self_play_buffer
 * (src/chess_zero/worker/self_play.py:113-154), ChessEnv.step/_game_over/_resign/adjudicate
 * (src/chess_zero/env/chess_env.py:79-129) and the python-chess calls those make (Board(), push, legal_moves, fen(),
 * result(claim_draw=True); python-chess==0.25.1, requirements.txt:1).
 *
 * Host only (libcaz_mcts.so, no CUDA): the leaves come out as the 68-byte board records of caz_encode
 * (include/caz_b200.h), so one round of the search for ALL games is
 *     caz_mcts_collect(m, records, cap, &n);  caz_eval_records(h, records, n, policy, value, NULL, 0);
 *     caz_mcts_apply(m, policy, value, n);
 * Plain C types, int status codes (0 = ok, 1 = bad argument or call order, 2 = out of memory / threads), caller-owned
 * buffers, no exception crosses the boundary.  One caller thread per caz_mcts object; the object runs its own workers.
 */
#ifndef CAZ_MCTS_H
#define CAZ_MCTS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct caz_mcts caz_mcts;

/* PlayConfig of the reference (configs/normal.py:28-43) */
typedef struct caz_mcts_config {
  int32_t simulation_num_per_move; /* 800  */
  int32_t search_threads;          /* 16: descents of one game in flight per round (virtual loss between them) */
  double c_puct;                   /* 1.5  */
  double noise_eps;                /* 0.25 */
  double dirichlet_alpha;          /* 0.3  */
  double tau_decay_rate;           /* 0.99 */
  int32_t virtual_loss;            /* 3    */
  int32_t resign_enabled;          /* 1    */
  double resign_threshold;         /* -0.8 */
  int32_t min_resign_turn;         /* 5    */
  int32_t max_game_length;         /* 1000 */
  int32_t n_labels;                /* 1968 */
  int32_t continuous;              /* 1: start a new game when one ends (SelfPlayWorker), 0: games stay finished */
  int32_t record_play_data;        /* 1: keep the [observation, policy, z] rows of finished games for caz_mcts_drain_play_data */
} caz_mcts_config;

typedef struct caz_mcts_stats {
  int64_t descents;         /* completed simulations (search_my_move calls from the root) */
  int64_t leaves_evaluated; /* positions sent to the network */
  int64_t terminal_hits;    /* descents that ended in a finished game (no evaluation) */
  int64_t withdrawn;        /* descents withdrawn because they ran into a leaf still being evaluated */
  int64_t moves_played, games_finished, white_wins, black_wins, draws, resignations, adjudications;
  int64_t tree_nodes_last;  /* nodes in the most recently completed move's tree */
} caz_mcts_stats;

/* move_label: int32 [64*64*5], (from*64+to)*5 + promotion (0 none, 1 n, 2 b, 3 r, 4 q) -> index into Config.labels
 * (config.py:76-123) or -1; unflipped_index: int32 [n_labels] (config.py:185).  Squares: a1 = 0 ... h8 = 63. */
int caz_mcts_create(const caz_mcts_config* cfg, int32_t n_games, uint64_t seed, const int32_t* move_label,
                    const int32_t* unflipped_index, int32_t n_threads, caz_mcts** out);
void caz_mcts_destroy(caz_mcts* m);

/* Walks the descents of every game (at most search_threads per game); records: uint8 [capacity][68] out,
 * capacity >= n_games * search_threads.  *n_leaves = 0 only when every game is finished (continuous = 0). */
int caz_mcts_collect(caz_mcts* m, uint8_t* records, int32_t capacity, int32_t* n_leaves);
/* policy: float32 [n][n_labels] and value: float32 [n] for the leaves of the last collect, in the same order. */
int caz_mcts_apply(caz_mcts* m, const float* policy, const float* value, int32_t n);

/* caz_mcts_collect that also lists every leaf's legal moves as policy labels (CSR: offsets int32 [n + 1], labels int32
 * [offsets[n]], edge order = python-chess move order, -1 for a move without a label; edge_capacity >= capacity * 256):
 * the input of caz_submit_records_priors, which then returns one prior per legal move instead of 1968 probabilities. */
int caz_mcts_collect_csr(caz_mcts* m, uint8_t* records, int32_t capacity, int32_t* n_leaves, int32_t* offsets,
                         int32_t* labels, int32_t edge_capacity);
/* caz_mcts_apply from priors already pushed on the device (select_action_q_and_u's p = P[a] / (1e-8 + sum), float64
 * [offsets[n]] in the order of the last caz_mcts_collect_csr): bit-identical trees to caz_mcts_apply. */
int caz_mcts_apply_priors(caz_mcts* m, const double* priors, const float* value, int32_t n);

int caz_mcts_get_stats(const caz_mcts* m, caz_mcts_stats* out);
/* ChessEnv.update(fen): restart game `game` from a position */
int caz_mcts_set_position(caz_mcts* m, int32_t game, const char* fen);
/* env.observation of a game (board.fen()); returns the length, or -1 if buf is too small */
int caz_mcts_game_fen(const caz_mcts* m, int32_t game, char* buf, int32_t cap);
/* 0 running, 1 white won, 2 black won, 3 draw (only meaningful with continuous = 0) */
int caz_mcts_game_result(const caz_mcts* m, int32_t game);
/* Root statistics of a game's current search: for each legal move (python-chess order) label, n, q, p. Returns the
 * number of edges written (<= cap), -1 if the root has not been expanded yet. */
int caz_mcts_root_edges(const caz_mcts* m, int32_t game, int32_t* label, int32_t* n, double* q, double* p, int32_t cap);

/* The training rows of the games finished since the last call (self_play_buffer's `data`, self_play.py:135-152;
 * what SelfPlayWorker.flush_buffer writes as JSON [[fen, policy[n_labels], z], ...], :90-100), serialised as
 *   fen bytes, 0 | z int8 (+1 the side that moved won, -1 lost, 0 draw) | n uint16 | n x (label uint16, probability float32)
 * per move, little endian, policy = visit counts / total over the labels visited.  Returns the number of bytes the
 * pending data needs; if that exceeds cap nothing is copied and nothing is dropped. */
int64_t caz_mcts_drain_play_data(caz_mcts* m, uint8_t* buf, int64_t cap);

/* Test hook: `draws` samples of the root's Dirichlet([alpha] * n) noise from the search's own generator (row major
 * [draws][n], float64), seeded like game 0 of a search created with `seed`. */
int caz_mcts_debug_dirichlet(uint64_t seed, double alpha, int32_t n, int32_t draws, double* out);

/* ---- the chess rules underneath (tests pin them with published perft counts) ---- */
uint64_t caz_chess_perft(const char* fen, int32_t depth);
/* legal moves in python-chess generation order, space separated UCI; returns the move count, -1 if buf too small */
int caz_chess_legal_moves(const char* fen, char* buf, int32_t cap);
/* Board(fen); push_uci each of `moves` (space separated); returns result(claim_draw=True): 0 "*", 1 "1-0", 2 "0-1",
 * 3 "1/2-1/2", -1 on an illegal move; fen_out (may be NULL) receives board.fen() afterwards */
int caz_chess_play(const char* fen, const char* moves, char* fen_out, int32_t cap);
/* the 68-byte record of caz_encode for a FEN, as this library derives it */
int caz_chess_record(const char* fen, uint8_t* record68);

#ifdef __cplusplus
}
#endif
#endif /* CAZ_MCTS_H */
