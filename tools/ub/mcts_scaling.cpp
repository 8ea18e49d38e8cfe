#include "chess/mcts.hpp"
#include <chrono>
#include <cstdio>
using namespace caz::chess;
int main(int argc, char** argv) {
  int G = 64;
  std::vector<int32_t> labels(64*64*5, -1), unflip(1968);
  int k = 0;
  for (int f = 0; f < 64; ++f) for (int t = 0; t < 64; ++t) for (int p = 0; p < 5; ++p) labels[(f*64+t)*5+p] = (k++) % 1968;
  for (int i = 0; i < 1968; ++i) unflip[i] = 1967 - i;
  std::vector<float> pol((size_t)G*16*1968), val(G*16, 0.01f);
  std::mt19937 r(1);
  for (auto& x : pol) x = (r() % 1000) / 1e6f;
  for (int threads : {1, 2, 4, 8, 16}) {
    MctsConfig c;
    Mcts m(c, G, 3, labels.data(), unflip.data(), threads);
    std::vector<uint8_t> rec((size_t)G*16*68);
    double tc = 0, ta = 0;
    auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 3) {
      int n;
      auto a = std::chrono::steady_clock::now();
      m.collect(rec.data(), G*16, &n);
      auto b = std::chrono::steady_clock::now();
      m.apply(pol.data(), val.data(), n);
      auto d = std::chrono::steady_clock::now();
      tc += std::chrono::duration<double>(b - a).count(); ta += std::chrono::duration<double>(d - b).count();
    }
    printf("%d threads: %.0f sims/s collect %.2f apply %.2f\n", threads, m.stats.descents / 3.0, tc, ta);
  }
}
