// BASELINE config 1 in the style of the reference's performance tests
// (apf/performance_tests/crossfade.cpp:105-142: fixed repetitions, one timer
// around the loop): apf::conv::StaticConvolver, 1 channel, 1024-sample block,
// 8192-tap IR, 48 kHz -- through the drop-in facade, i.e. on the GPU.
// Wall time via steady_clock (apf::StopWatch measures CPU time).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "apf/convolver.h"

int main(int argc, char* argv[])
{
  const size_t block_size = 1024, taps = 8192;
  const int repetitions = argc > 1 ? std::atoi(argv[1]) : 2000;
  std::vector<float> ir(taps), signal(block_size);
  unsigned s = 12345u;
  auto rnd = [&s] { s = s * 1664525u + 1013904223u; return float(s >> 8) / 8388608.0f - 1.0f; };
  for (auto& v : ir) v = rnd() * 0.01f;
  for (auto& v : signal) v = rnd();

  apf::conv::StaticConvolver conv(block_size, ir.begin(), ir.end());
  float sink = 0.0f;
  for (int i = 0; i < 100; ++i) { conv.add_block(signal.begin()); sink += conv.convolve()[0]; }
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < repetitions; ++i)
  {
    conv.add_block(signal.begin());
    sink += conv.convolve()[0];
  }
  auto t1 = std::chrono::steady_clock::now();
  const double secs = std::chrono::duration<double>(t1 - t0).count();
  const double us = secs / repetitions * 1e6;
  std::printf("{\"workload\": \"cfg1 StaticConvolver 1ch B=1024 8192 taps\", \"repetitions\": %d, "
      "\"us_per_block\": %.2f, \"x_realtime\": %.1f, \"sink\": %g}\n",
      repetitions, us, (double(block_size) / 48000.0) / (secs / repetitions), double(sink));
  return 0;
}
