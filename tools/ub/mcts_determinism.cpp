// Behaviour check for search refactors: 3000 rounds of 16 games against a fixed fake evaluator; the hash of every board
// record requested and the counters must not change (g++ -O2 -std=c++17 -I chess-alpha-zero_b200/csrc tools/ub/mcts_determinism.cpp -lpthread).
#include "chess/mcts.hpp"
#include <cstdio>
using namespace caz::chess;
int main() {
  int G = 16;
  std::vector<int32_t> labels(64*64*5, -1), unflip(1968);
  int k = 0;
  for (int f = 0; f < 64; ++f) for (int t = 0; t < 64; ++t) for (int p = 0; p < 5; ++p) labels[(f*64+t)*5+p] = (k++) % 1968;
  for (int i = 0; i < 1968; ++i) unflip[i] = 1967 - i;
  std::vector<float> pol((size_t)G*16*1968), val(G*16, 0.01f);
  std::mt19937 r(1);
  for (auto& x : pol) x = (r() % 1000) / 1e6f;
  for (size_t i = 0; i < val.size(); ++i) val[i] = ((int)(r() % 200) - 100) / 150.0f;
  MctsConfig c;
  c.simulation_num_per_move = 100;
  Mcts m(c, G, 3, labels.data(), unflip.data(), 2);
  std::vector<uint8_t> rec((size_t)G*16*68);
  uint64_t h = 1469598103934665603ull;
  for (int round = 0; round < 3000; ++round) {
    int n;
    m.collect(rec.data(), G*16, &n);
    for (int i = 0; i < n * 68; ++i) { h ^= rec[i]; h *= 1099511628211ull; }
    m.apply(pol.data(), val.data(), n);
  }
  printf("hash %016llx descents %lld moves %lld games %lld terminal %lld\n", (unsigned long long)h, (long long)m.stats.descents,
         (long long)m.stats.moves_played, (long long)m.stats.games_finished, (long long)m.stats.terminal_hits);
}
