// Host-side chess rules for the batched MCTS driver: the part of python-chess 0.25.1 (a pinned
// dependency of the reference, requirements.txt:1, NOT vendored in /root/reference) that the reference's search
// touches -- chess_env.py:50,62,91-106 (Board(), push_uci, result(claim_draw=True)), player_chess.py:183,204,263-275
// (fen(), legal_moves iteration order).  Restated from python-chess's published behaviour:
//   * squares a1 = 0 .. h8 = 63; legal moves are produced in python-chess's generation order (non-pawn pieces from the
//     highest square down, targets from the highest square down; castling, h-side first; pawn captures; single pushes;
//     double pushes; en passant; when in check: king moves first, then blocks/captures of a single checker) because
//     select_action_q_and_u breaks score ties by that order (player_chess.py:283-297);
//   * fen() writes the en-passant square only when an en-passant capture is legal (python-chess's default
//     en_passant="legal"), and that string is both the tree key and the encoder's input;
//   * result(claim_draw=True): checkmate, else fifty-move / threefold claim (incl. "a legal move would repeat"),
//     else insufficient material, else stalemate.
// Parity with python-chess itself is UNPINNED here (the package is not installed); tests pin the move generator with
// the published perft counts and the FEN/draw logic with hand-checked positions.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace caz {
namespace chess {

typedef uint64_t BB;
enum Piece : int { PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5, NO_PIECE = 6 };
enum Color : int { BLACK = 0, WHITE = 1 };   // python-chess: chess.WHITE = True
enum : int { CR_WK = 1, CR_WQ = 2, CR_BK = 4, CR_BQ = 8 };

// move = from | to << 6 | promotion piece << 12 (0 = none, else KNIGHT..QUEEN)
typedef uint16_t Move;
inline Move make_move(int from, int to, int promo = 0) { return static_cast<Move>(from | (to << 6) | (promo << 12)); }
inline int move_from(Move m) { return m & 63; }
inline int move_to(Move m) { return (m >> 6) & 63; }
inline int move_promo(Move m) { return (m >> 12) & 7; }

inline BB sq_bb(int s) { return BB(1) << s; }
inline int lsb(BB b) { return __builtin_ctzll(b); }
inline int msb(BB b) { return 63 - __builtin_clzll(b); }
inline int popcount(BB b) { return __builtin_popcountll(b); }
inline int file_of(int s) { return s & 7; }
inline int rank_of(int s) { return s >> 3; }

struct Tables {
  BB knight[64], king[64], pawn_att[2][64];
  BB ray[8][64];        // N, S, E, W, NE, SW, NW, SE: d ^ 1 is the opposite direction, even d points to higher squares
  BB between[64][64];   // squares strictly between two aligned squares
  BB line[64][64];      // the whole line through two aligned squares (incl. both), 0 if not aligned
  uint64_t z_piece[2][6][64], z_castle[16], z_ep[8], z_turn;
  Tables() {
    static const int dr[8] = {1, -1, 0, 0, 1, -1, 1, -1}, df[8] = {0, 0, 1, -1, 1, -1, -1, 1};
    for (int s = 0; s < 64; ++s) {
      const int r = rank_of(s), f = file_of(s);
      knight[s] = king[s] = pawn_att[0][s] = pawn_att[1][s] = 0;
      static const int kr[8] = {2, 2, -2, -2, 1, 1, -1, -1}, kf[8] = {1, -1, 1, -1, 2, -2, 2, -2};
      for (int i = 0; i < 8; ++i) {
        int rr = r + kr[i], ff = f + kf[i];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) knight[s] |= sq_bb(rr * 8 + ff);
        rr = r + dr[i]; ff = f + df[i];
        if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) king[s] |= sq_bb(rr * 8 + ff);
      }
      for (int d = -1; d <= 1; d += 2) {
        if (f + d >= 0 && f + d < 8) {
          if (r + 1 < 8) pawn_att[WHITE][s] |= sq_bb((r + 1) * 8 + f + d);
          if (r - 1 >= 0) pawn_att[BLACK][s] |= sq_bb((r - 1) * 8 + f + d);
        }
      }
      for (int d = 0; d < 8; ++d) {
        ray[d][s] = 0;
        for (int rr = r + dr[d], ff = f + df[d]; rr >= 0 && rr < 8 && ff >= 0 && ff < 8; rr += dr[d], ff += df[d])
          ray[d][s] |= sq_bb(rr * 8 + ff);
      }
    }
    for (int a = 0; a < 64; ++a)
      for (int b = 0; b < 64; ++b) {
        between[a][b] = line[a][b] = 0;
        for (int d = 0; d < 8; ++d)
          if (ray[d][a] & sq_bb(b)) {
            between[a][b] = ray[d][a] & ray[d ^ 1][b];
            line[a][b] = ray[d][a] | ray[d ^ 1][a] | sq_bb(a);
          }
      }
    uint64_t x = 0x9E3779B97F4A7C15ull;   // splitmix64
    auto next = [&x]() {
      uint64_t z = (x += 0x9E3779B97F4A7C15ull);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
    };
    for (int c = 0; c < 2; ++c)
      for (int p = 0; p < 6; ++p)
        for (int s = 0; s < 64; ++s) z_piece[c][p][s] = next();
    for (int i = 0; i < 16; ++i) z_castle[i] = next();
    for (int i = 0; i < 8; ++i) z_ep[i] = next();
    z_turn = next();
  }
};
inline const Tables& T() {
  static const Tables t;
  return t;
}

// first blocker along a ray: the lowest set bit for rays towards higher squares (even d), else the highest
inline BB ray_attacks(int d, int s, BB occ) {
  const BB r = T().ray[d][s], blockers = r & occ;
  if (!blockers) return r;
  const bool positive = (d & 1) == 0;
  const int b = positive ? lsb(blockers) : msb(blockers);
  return r ^ T().ray[d][b];
}
inline BB rook_attacks(int s, BB occ) {
  return ray_attacks(0, s, occ) | ray_attacks(1, s, occ) | ray_attacks(2, s, occ) | ray_attacks(3, s, occ);
}
inline BB bishop_attacks(int s, BB occ) {
  return ray_attacks(4, s, occ) | ray_attacks(5, s, occ) | ray_attacks(6, s, occ) | ray_attacks(7, s, occ);
}

struct Position {
  BB by_piece[6];
  BB by_color[2];
  int turn;        // WHITE / BLACK
  int castling;    // CR_* bits
  int ep_square;   // set after every double pawn push (python-chess Board.ep_square), -1 otherwise
  int halfmove, fullmove;

  BB occupied() const { return by_color[0] | by_color[1]; }
  int piece_at(int s) const {
    const BB b = sq_bb(s);
    for (int p = 0; p < 6; ++p)
      if (by_piece[p] & b) return p;
    return NO_PIECE;
  }
  int king_sq(int c) const { return lsb(by_piece[KING] & by_color[c]); }

  BB attackers(int color, int s, BB occ) const {
    const Tables& t = T();
    const BB rq = by_piece[ROOK] | by_piece[QUEEN], bq = by_piece[BISHOP] | by_piece[QUEEN];
    return ((t.king[s] & by_piece[KING]) | (t.knight[s] & by_piece[KNIGHT]) | (rook_attacks(s, occ) & rq) |
            (bishop_attacks(s, occ) & bq) | (t.pawn_att[color ^ 1][s] & by_piece[PAWN])) &
           by_color[color] & occ;
  }
  bool attacked_any(int color, BB squares, BB occ) const {
    while (squares) {
      const int s = lsb(squares);
      squares &= squares - 1;
      if (attackers(color, s, occ)) return true;
    }
    return false;
  }
  bool in_check() const { return attackers(turn ^ 1, king_sq(turn), occupied()) != 0; }

  void set_startpos() { from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"); }

  bool from_fen(const std::string& fen) {
    std::memset(by_piece, 0, sizeof(by_piece));
    by_color[0] = by_color[1] = 0;
    size_t i = 0;
    int r = 7, f = 0;
    for (; i < fen.size() && fen[i] != ' '; ++i) {
      const char c = fen[i];
      if (c == '/') { --r; f = 0; continue; }
      if (c >= '1' && c <= '8') { f += c - '0'; continue; }
      static const char* names = "pnbrqk";
      const char lc = static_cast<char>(c | 0x20);
      const char* p = std::strchr(names, lc);
      if (!p || r < 0 || f > 7) return false;
      const int s = r * 8 + f++;
      by_piece[p - names] |= sq_bb(s);
      by_color[c == lc ? BLACK : WHITE] |= sq_bb(s);
    }
    std::vector<std::string> fields;
    std::string cur;
    for (++i; i <= fen.size(); ++i) {
      if (i == fen.size() || fen[i] == ' ') { if (!cur.empty()) fields.push_back(cur); cur.clear(); }
      else cur.push_back(fen[i]);
    }
    turn = (fields.size() > 0 && fields[0] == "b") ? BLACK : WHITE;
    castling = 0;
    if (fields.size() > 1)
      for (char c : fields[1]) {
        if (c == 'K') castling |= CR_WK; else if (c == 'Q') castling |= CR_WQ;
        else if (c == 'k') castling |= CR_BK; else if (c == 'q') castling |= CR_BQ;
      }
    // python-chess clean_castling_rights(): a right needs king and rook on their original squares
    const BB wk = by_piece[KING] & by_color[WHITE], bk = by_piece[KING] & by_color[BLACK];
    const BB wr = by_piece[ROOK] & by_color[WHITE], br = by_piece[ROOK] & by_color[BLACK];
    if (!(wk & sq_bb(4)) || !(wr & sq_bb(7))) castling &= ~CR_WK;
    if (!(wk & sq_bb(4)) || !(wr & sq_bb(0))) castling &= ~CR_WQ;
    if (!(bk & sq_bb(60)) || !(br & sq_bb(63))) castling &= ~CR_BK;
    if (!(bk & sq_bb(60)) || !(br & sq_bb(56))) castling &= ~CR_BQ;
    ep_square = -1;
    if (fields.size() > 2 && fields[2].size() == 2 && fields[2][0] >= 'a' && fields[2][0] <= 'h')
      ep_square = (fields[2][1] - '1') * 8 + (fields[2][0] - 'a');
    halfmove = fields.size() > 3 ? std::atoi(fields[3].c_str()) : 0;
    fullmove = fields.size() > 4 ? std::atoi(fields[4].c_str()) : 1;
    return popcount(wk) == 1 && popcount(bk) == 1;
  }

  // ---- move generation, python-chess order ------------------------------------------------------------------
  // pseudo-legal moves of the pieces in from_mask to the squares in to_mask (generate_pseudo_legal_moves)
  int gen_pseudo(BB from_mask, BB to_mask, Move* out) const {
    const Tables& t = T();
    int n = 0;
    const BB ours = by_color[turn], theirs = by_color[turn ^ 1], occ = ours | theirs;
    BB non_pawns = ours & ~by_piece[PAWN] & from_mask;
    while (non_pawns) {
      const int from = msb(non_pawns);
      non_pawns ^= sq_bb(from);
      const BB b = sq_bb(from);
      BB att;
      if (by_piece[KNIGHT] & b) att = t.knight[from];
      else if (by_piece[KING] & b) att = t.king[from];
      else {
        att = 0;
        if ((by_piece[BISHOP] | by_piece[QUEEN]) & b) att |= bishop_attacks(from, occ);
        if ((by_piece[ROOK] | by_piece[QUEEN]) & b) att |= rook_attacks(from, occ);
      }
      att &= ~ours & to_mask;
      while (att) {
        const int to = msb(att);
        att ^= sq_bb(to);
        out[n++] = make_move(from, to);
      }
    }
    if (from_mask & by_piece[KING]) n += gen_castling(to_mask, out + n);
    const BB pawns = by_piece[PAWN] & ours & from_mask;
    if (!pawns) return n;
    BB capturers = pawns;
    while (capturers) {
      const int from = msb(capturers);
      capturers ^= sq_bb(from);
      BB targets = t.pawn_att[turn][from] & theirs & to_mask;
      while (targets) {
        const int to = msb(targets);
        targets ^= sq_bb(to);
        if (rank_of(to) == 0 || rank_of(to) == 7) {
          out[n++] = make_move(from, to, QUEEN); out[n++] = make_move(from, to, ROOK);
          out[n++] = make_move(from, to, BISHOP); out[n++] = make_move(from, to, KNIGHT);
        } else {
          out[n++] = make_move(from, to);
        }
      }
    }
    BB single, dbl;
    if (turn == WHITE) {
      single = (pawns << 8) & ~occ;
      dbl = (single << 8) & ~occ & 0x00000000FF000000ull;   // rank 4 (BB_RANK_3 | BB_RANK_4 with single on rank 3)
    } else {
      single = (pawns >> 8) & ~occ;
      dbl = (single >> 8) & ~occ & 0x000000FF00000000ull;   // rank 5
    }
    single &= to_mask;
    dbl &= to_mask;
    while (single) {
      const int to = msb(single);
      single ^= sq_bb(to);
      const int from = to + (turn == BLACK ? 8 : -8);
      if (rank_of(to) == 0 || rank_of(to) == 7) {
        out[n++] = make_move(from, to, QUEEN); out[n++] = make_move(from, to, ROOK);
        out[n++] = make_move(from, to, BISHOP); out[n++] = make_move(from, to, KNIGHT);
      } else {
        out[n++] = make_move(from, to);
      }
    }
    while (dbl) {
      const int to = msb(dbl);
      dbl ^= sq_bb(to);
      out[n++] = make_move(to + (turn == BLACK ? 16 : -16), to);
    }
    if (ep_square >= 0) n += gen_ep(from_mask, to_mask, out + n);
    return n;
  }

  int gen_ep(BB from_mask, BB to_mask, Move* out) const {
    if (ep_square < 0 || !(sq_bb(ep_square) & to_mask) || (sq_bb(ep_square) & occupied())) return 0;
    const BB rank = turn == WHITE ? 0x000000FF00000000ull : 0x00000000FF000000ull;   // rank 5 / rank 4
    BB capturers = by_piece[PAWN] & by_color[turn] & from_mask & T().pawn_att[turn ^ 1][ep_square] & rank;
    int n = 0;
    while (capturers) {
      const int from = msb(capturers);
      capturers ^= sq_bb(from);
      out[n++] = make_move(from, ep_square);
    }
    return n;
  }

  int gen_castling(BB to_mask, Move* out) const {
    const int k = turn == WHITE ? 4 : 60;
    const BB occ = occupied();
    if (!(by_piece[KING] & by_color[turn] & sq_bb(k)) || attackers(turn ^ 1, k, occ)) return 0;
    int n = 0;
    const int them = turn ^ 1;
    // candidates are the rook squares, highest first: h-side before a-side
    for (int side = 0; side < 2; ++side) {
      const int right = turn == WHITE ? (side == 0 ? CR_WK : CR_WQ) : (side == 0 ? CR_BK : CR_BQ);
      if (!(castling & right)) continue;
      const int rook = side == 0 ? k + 3 : k - 4;
      if (!(sq_bb(rook) & to_mask)) continue;   // python-chess masks the rook square (its internal king-takes-rook form)
      const int king_to = side == 0 ? k + 2 : k - 2, rook_to = side == 0 ? k + 1 : k - 1;
      const BB king_path = T().between[k][king_to], rook_path = T().between[rook][rook_to];
      const BB without = occ ^ sq_bb(k) ^ sq_bb(rook);
      if (without & (king_path | rook_path | sq_bb(king_to) | sq_bb(rook_to))) continue;
      if (attacked_any(them, king_path | sq_bb(k), occ ^ sq_bb(k))) continue;
      if (attacked_any(them, sq_bb(king_to), without ^ sq_bb(rook_to))) continue;
      out[n++] = make_move(k, king_to);
    }
    return n;
  }

  bool is_castling(Move m) const {
    const int from = move_from(m), to = move_to(m);
    return (by_piece[KING] & sq_bb(from)) && (to - from == 2 || from - to == 2);
  }
  bool is_en_passant(Move m) const {
    const int from = move_from(m), to = move_to(m);
    return to == ep_square && (by_piece[PAWN] & sq_bb(from)) && file_of(from) != file_of(to) && !(occupied() & sq_bb(to));
  }
  bool is_capture(Move m) const { return (by_color[turn ^ 1] & sq_bb(move_to(m))) || is_en_passant(m); }
  bool is_zeroing(Move m) const { return (by_piece[PAWN] & sq_bb(move_from(m))) || is_capture(m); }
  bool reduces_castling_rights(Move m) const {
    if (!castling) return false;
    Position c = *this;
    c.push(m);
    return c.castling != castling;
  }
  bool is_irreversible(Move m) const { return is_zeroing(m) || reduces_castling_rights(m); }

  BB slider_blockers(int king) const {
    const BB rq = by_piece[ROOK] | by_piece[QUEEN], bq = by_piece[BISHOP] | by_piece[QUEEN];
    BB snipers = ((rook_attacks(king, 0) & rq) | (bishop_attacks(king, 0) & bq)) & by_color[turn ^ 1];
    const BB occ = occupied();
    BB blockers = 0;
    while (snipers) {
      const int s = msb(snipers);
      snipers ^= sq_bb(s);
      const BB b = T().between[king][s] & occ;
      if (b && !(b & (b - 1))) blockers |= b;
    }
    return blockers & by_color[turn];
  }

  bool ep_is_safe(int king, Move m) const {
    // remove both pawns, put ours on the ep square: is the king attacked by a slider (or anything else)?
    const int from = move_from(m), to = move_to(m), cap = to + (turn == WHITE ? -8 : 8);
    const BB occ = (occupied() ^ sq_bb(from) ^ sq_bb(cap)) | sq_bb(to);
    Position c = *this;
    c.by_piece[PAWN] &= ~sq_bb(cap);
    c.by_color[turn ^ 1] &= ~sq_bb(cap);
    return c.attackers(turn ^ 1, king, occ) == 0;
  }
  bool is_safe(int king, BB blockers, Move m) const {
    const int from = move_from(m), to = move_to(m);
    if (from == king) {
      if (is_castling(m)) return true;
      return attackers(turn ^ 1, to, occupied()) == 0;
    }
    if (is_en_passant(m)) return ep_is_safe(king, m);
    return !(blockers & sq_bb(from)) || (T().line[from][to] & sq_bb(king));
  }

  // legal moves in python-chess's order (generate_legal_moves); returns the count (<= 218)
  int legal_moves(Move* out) const {
    Move tmp[256];
    int nt = 0;
    const int king = king_sq(turn);
    const BB blockers = slider_blockers(king);
    const BB occ = occupied();
    const BB checkers = attackers(turn ^ 1, king, occ);
    if (checkers) {
      // _generate_evasions
      const BB sliders = checkers & (by_piece[BISHOP] | by_piece[ROOK] | by_piece[QUEEN]);
      BB attacked = 0, s = sliders;
      while (s) {
        const int c = msb(s);
        s ^= sq_bb(c);
        attacked |= T().line[king][c] & ~sq_bb(c);
      }
      BB kt = T().king[king] & ~by_color[turn] & ~attacked;
      while (kt) {
        const int to = msb(kt);
        kt ^= sq_bb(to);
        tmp[nt++] = make_move(king, to);
      }
      const int checker = msb(checkers);
      if (sq_bb(checker) == checkers) {
        const BB target = T().between[king][checker] | checkers;
        nt += gen_pseudo(~by_piece[KING], target, tmp + nt);
        if (ep_square >= 0 && !(sq_bb(ep_square) & target)) {
          const int last_double = ep_square + (turn == WHITE ? -8 : 8);
          if (last_double == checker) nt += gen_ep(~BB(0), ~BB(0), tmp + nt);
        }
      }
    } else {
      nt = gen_pseudo(~BB(0), ~BB(0), tmp);
    }
    int n = 0;
    for (int i = 0; i < nt; ++i)
      if (is_safe(king, blockers, tmp[i])) out[n++] = tmp[i];
    return n;
  }
  bool has_legal_move() const {
    Move m[256];
    return legal_moves(m) > 0;
  }
  bool has_legal_en_passant() const {
    if (ep_square < 0) return false;
    Move m[4];
    const int n = gen_ep(~BB(0), ~BB(0), m);
    if (!n) return false;
    const int king = king_sq(turn);
    // in check the capture must also resolve the check: test fully
    for (int i = 0; i < n; ++i) {
      Position c = *this;
      c.push(m[i]);
      if (!c.attackers(c.turn, king, c.occupied())) return true;   // our king (same square) not attacked after the move
    }
    return false;
  }

  // Board.push for a (pseudo-)legal move
  void push(Move m) {
    const int from = move_from(m), to = move_to(m), promo = move_promo(m);
    const BB fb = sq_bb(from), tb = sq_bb(to);
    const int us = turn, them = turn ^ 1;
    const int piece = piece_at(from);
    const bool ep = is_en_passant(m), castle = is_castling(m);
    const bool zero = piece == PAWN || (by_color[them] & tb) || ep;
    halfmove = zero ? 0 : halfmove + 1;
    if (us == BLACK) ++fullmove;
    // captures
    if (ep) {
      const int cap = to + (us == WHITE ? -8 : 8);
      by_piece[PAWN] &= ~sq_bb(cap);
      by_color[them] &= ~sq_bb(cap);
    } else if (by_color[them] & tb) {
      for (int p = 0; p < 6; ++p) by_piece[p] &= ~tb;
      by_color[them] &= ~tb;
    }
    ep_square = -1;
    if (piece == PAWN && (to - from == 16 || from - to == 16)) ep_square = (from + to) / 2;
    // castling rights
    if (piece == KING) castling &= us == WHITE ? ~(CR_WK | CR_WQ) : ~(CR_BK | CR_BQ);
    if (from == 7 || to == 7) castling &= ~CR_WK;
    if (from == 0 || to == 0) castling &= ~CR_WQ;
    if (from == 63 || to == 63) castling &= ~CR_BK;
    if (from == 56 || to == 56) castling &= ~CR_BQ;
    // move the piece
    by_piece[piece] &= ~fb;
    by_color[us] &= ~fb;
    by_piece[promo ? promo : piece] |= tb;
    by_color[us] |= tb;
    if (castle) {
      const int rook_from = to > from ? from + 3 : from - 4, rook_to = to > from ? from + 1 : from - 1;
      by_piece[ROOK] = (by_piece[ROOK] & ~sq_bb(rook_from)) | sq_bb(rook_to);
      by_color[us] = (by_color[us] & ~sq_bb(rook_from)) | sq_bb(rook_to);
    }
    turn = them;
  }

  // Board._transposition_key(): pieces, turn, castling rights, en-passant square if a capture is legal
  uint64_t transposition_key() const {
    const Tables& t = T();
    uint64_t k = 0;
    for (int c = 0; c < 2; ++c)
      for (int p = 0; p < 6; ++p) {
        BB b = by_piece[p] & by_color[c];
        while (b) {
          const int s = lsb(b);
          b &= b - 1;
          k ^= t.z_piece[c][p][s];
        }
      }
    k ^= t.z_castle[castling];
    if (turn == BLACK) k ^= t.z_turn;
    if (has_legal_en_passant()) k ^= t.z_ep[file_of(ep_square)] ^ 0x5bd1e995u;
    return k;
  }
  // the reference's tree key, state_key(env) = fen() without the fullmove number (player_chess.py:391-398):
  // everything in the transposition key plus the halfmove clock
  uint64_t state_key() const { return state_key_of(transposition_key(), halfmove); }
  // (the search keeps the transposition key of the current position in its move stack: no second pass over the board)
  static uint64_t state_key_of(uint64_t transposition, int halfmove_clock) {
    uint64_t k = transposition ^ (0x9E3779B97F4A7C15ull * static_cast<uint64_t>(halfmove_clock + 1));
    k ^= k >> 29;
    return k * 0xBF58476D1CE4E5B9ull;
  }

  bool insufficient_material_for(int color) const {
    const BB own = by_color[color];
    if (own & (by_piece[PAWN] | by_piece[ROOK] | by_piece[QUEEN])) return false;
    if (own & by_piece[KNIGHT])
      return popcount(own) <= 2 && !(by_color[color ^ 1] & ~by_piece[KING] & ~by_piece[QUEEN]);
    if (own & by_piece[BISHOP]) {
      const BB dark = 0xAA55AA55AA55AA55ull, light = ~dark;
      const bool same_color = !(by_piece[BISHOP] & dark) || !(by_piece[BISHOP] & light);
      return same_color && !by_piece[PAWN] && !by_piece[KNIGHT];
    }
    return true;
  }
  bool insufficient_material() const { return insufficient_material_for(WHITE) && insufficient_material_for(BLACK); }

  std::string fen() const {
    static const char* names = "pnbrqk";
    std::string s;
    for (int r = 7; r >= 0; --r) {
      int empty = 0;
      for (int f = 0; f < 8; ++f) {
        const int sq = r * 8 + f, p = piece_at(sq);
        if (p == NO_PIECE) { ++empty; continue; }
        if (empty) { s.push_back(static_cast<char>('0' + empty)); empty = 0; }
        const char c = names[p];
        s.push_back((by_color[WHITE] & sq_bb(sq)) ? static_cast<char>(c - 32) : c);
      }
      if (empty) s.push_back(static_cast<char>('0' + empty));
      if (r) s.push_back('/');
    }
    s += turn == WHITE ? " w " : " b ";
    std::string c;
    if (castling & CR_WK) c.push_back('K');
    if (castling & CR_WQ) c.push_back('Q');
    if (castling & CR_BK) c.push_back('k');
    if (castling & CR_BQ) c.push_back('q');
    s += c.empty() ? "-" : c;
    s.push_back(' ');
    if (has_legal_en_passant()) {
      s.push_back(static_cast<char>('a' + file_of(ep_square)));
      s.push_back(static_cast<char>('1' + rank_of(ep_square)));
    } else {
      s.push_back('-');
    }
    s += " " + std::to_string(halfmove) + " " + std::to_string(fullmove);
    return s;
  }

  // the 68-byte board record of caz_encode (include/caz_b200.h): the position as fen() writes it
  void record(uint8_t* rec) const {
    static const int code_of[6] = {6, 5, 4, 3, 2, 1};   // "KQRBNPkqrbnp": K=1 Q=2 R=3 B=4 N=5 P=6, black +6
    std::memset(rec, 0, 64);   // byte (7 - rank) * 8 + file = square ^ 56; one store per piece on the board
    for (int c = 0; c < 2; ++c)
      for (int p = 0; p < 6; ++p) {
        BB b = by_piece[p] & by_color[c];
        const uint8_t code = static_cast<uint8_t>(code_of[p] + (c == BLACK ? 6 : 0));
        while (b) {
          rec[lsb(b) ^ 56] = code;
          b &= b - 1;
        }
      }
    rec[64] = static_cast<uint8_t>((turn == BLACK ? 1 : 0) | ((castling & CR_WK) ? 2 : 0) | ((castling & CR_WQ) ? 4 : 0) |
                                   ((castling & CR_BK) ? 8 : 0) | ((castling & CR_BQ) ? 16 : 0));
    rec[65] = has_legal_en_passant() ? static_cast<uint8_t>((7 - rank_of(ep_square)) * 8 + file_of(ep_square)) : 255;
    rec[66] = static_cast<uint8_t>(halfmove & 255);
    rec[67] = static_cast<uint8_t>((halfmove >> 8) & 255);
  }
};

inline std::string move_uci(Move m) {
  std::string s;
  s.push_back(static_cast<char>('a' + file_of(move_from(m))));
  s.push_back(static_cast<char>('1' + rank_of(move_from(m))));
  s.push_back(static_cast<char>('a' + file_of(move_to(m))));
  s.push_back(static_cast<char>('1' + rank_of(move_to(m))));
  if (move_promo(m)) s.push_back("?nbrq"[move_promo(m)]);
  return s;
}
inline Move move_from_uci(const std::string& s) {
  if (s.size() < 4) return 0;
  int promo = 0;
  if (s.size() > 4) {
    const char* p = std::strchr("?nbrq", s[4]);
    promo = p ? static_cast<int>(p - "?nbrq") : 0;
  }
  return make_move((s[1] - '1') * 8 + (s[0] - 'a'), (s[3] - '1') * 8 + (s[2] - 'a'), promo);
}

// A game: the position plus what result(claim_draw=True) needs from the move stack.
struct Game {
  Position pos;
  // key of a position; was the move INTO it irreversible; how often the position has occurred since the last
  // irreversible move (this occurrence included); how many entries since then were repeats (occ >= 2)
  struct Hist { uint64_t key; bool irreversible_in; uint16_t occ; uint16_t repeats; };
  std::vector<Hist> hist;
  int num_halfmoves = 0;   // ChessEnv.num_halfmoves

  void reset(const Position& p) {
    pos = p;
    hist.clear();
    hist.push_back({pos.transposition_key(), true, 1, 0});
    num_halfmoves = 0;
  }
  void push(Move m) {
    const bool irr = pos.is_irreversible(m);
    pos.push(m);
    const uint64_t key = pos.transposition_key();
    uint16_t occ = 1, repeats = 0;
    if (!irr) {
      // positions since the last irreversible move, newest first (the entry that move led INTO is the oldest)
      for (size_t i = hist.size(); i-- > 0;) {
        occ += hist[i].key == key;
        if (hist[i].irreversible_in) break;
      }
      repeats = static_cast<uint16_t>(hist.back().repeats + (occ >= 2 ? 1 : 0));
    }
    hist.push_back({key, irr, occ, repeats});
    ++num_halfmoves;
  }
  void pop_to(size_t hist_size, const Position& p, int halfmoves) {   // undo a simulation
    hist.resize(hist_size);
    pos = p;
    num_halfmoves = halfmoves;
  }
  uint64_t state_key() const { return Position::state_key_of(hist.back().key, pos.halfmove); }

  // Board.can_claim_threefold_repetition(); move_at(i) yields the i-th legal move of the current position
  template <class MoveAt>
  bool can_claim_threefold(MoveAt move_at, int n_legal) const {
    if (hist.back().occ >= 3) return true;
    if (hist.back().repeats == 0) return false;   // no position occurred twice: no move can reach a third occurrence
    size_t first = hist.size() - 1;
    while (first > 0 && !hist[first].irreversible_in) --first;
    for (int m = 0; m < n_legal; ++m) {
      Position c = pos;
      c.push(move_at(m));
      const uint64_t k = c.transposition_key();
      int cnt = 0;
      for (size_t i = first; i < hist.size(); ++i) cnt += hist[i].key == k;
      if (cnt >= 2) return true;
    }
    return false;
  }

  // Board.result(claim_draw=True): 0 = "*", 1 = "1-0", 2 = "0-1", 3 = "1/2-1/2"
  template <class MoveAt>
  int result_with(MoveAt move_at, int n_legal) const {
    if (n_legal == 0 && pos.in_check()) return pos.turn == WHITE ? 2 : 1;
    if (pos.halfmove >= 100 && n_legal > 0) return 3;
    if (can_claim_threefold(move_at, n_legal)) return 3;
    if (pos.insufficient_material()) return 3;
    if (n_legal == 0) return 3;
    return 0;
  }
  int result(const Move* legal, int n_legal) const {
    return result_with([legal](int i) { return legal[i]; }, n_legal);
  }
  int result() const {
    Move m[256];
    const int n = pos.legal_moves(m);
    return result(m, n);
  }
};

inline uint64_t perft(const Position& p, int depth) {
  Move m[256];
  const int n = p.legal_moves(m);
  if (depth <= 1) return depth == 1 ? static_cast<uint64_t>(n) : 1;
  uint64_t total = 0;
  for (int i = 0; i < n; ++i) {
    Position c = p;
    c.push(m[i]);
    total += perft(c, depth - 1);
  }
  return total;
}

}  // namespace chess
}  // namespace caz
