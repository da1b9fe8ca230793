// C++ test of the apf::conv drop-in facade (ssr-pre-0.4.0_b200/cpp/apf/convolver.h)
// on a real GPU.  The cases restate what the reference's own unit test checks
// (apf/unit_tests/test_convolver.cpp:37-238: block size 8, 16-tap filter with
// 5, 4, 3 at taps 10..12, ramp signal; sections "silence", "impulse", "... and
// more", "StaticOutput impulse", "StaticOutput frequency domain", "combinations")
// with the same comparison rule as Catch's Approx (catch.hpp:2087-2088,2131).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <vector>

#include "apf/convolver.h"

namespace c = apf::conv;

static int failures = 0;
static int checks = 0;

static bool approx(float a, float b)
{
  const double eps = 100.0 * FLT_EPSILON;
  return std::fabs(double(a) - double(b)) < eps * (1.0 + std::fmax(std::fabs(a), std::fabs(b)));
}

#define CHECK(cond) do { ++checks; if (!(cond)) { ++failures; \
  std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)

#define CHECK_RANGE(left, right, range) \
  for (int i_ = 0; i_ < (range); ++i_) { ++checks; if (!approx((left)[i_], (right)[i_])) { ++failures; \
    std::printf("FAILED %s:%d: [%d] %g != %g\n", __FILE__, __LINE__, i_, double((left)[i_]), double((right)[i_])); } }

static float test_signal[] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.1f, 0.2f, 0.3f, 0.4f
  , 0.5f, 0.6f, 0.7f, 0.8f, 0.9f, 0.0f, 0.0f, 0.0f, 0.0f };
static float zeros[20] = { 0.0f };

struct Fixture
{
  Fixture()
    : partitions(c::min_partitions(8, 16))
    , conv_input(8, partitions)
    , conv_output(conv_input)
    , filter(8, filter_data(), filter_data() + 16)
  {}
  static float* filter_data()
  {
    static float d[16] = { 0.0f };
    d[10] = 5.0f; d[11] = 4.0f; d[12] = 3.0f;
    return d;
  }
  size_t partitions;
  c::Input conv_input;
  c::Output conv_output;
  c::Filter filter;
};

static void section_silence()
{
  Fixture f;
  f.conv_input.add_block(test_signal);
  float* result = f.conv_output.convolve();
  CHECK_RANGE(result, zeros, 8);
  f.conv_input.add_block(test_signal);
  result = f.conv_output.convolve();
  CHECK_RANGE(result, zeros, 8);
}

static void section_impulse()
{
  Fixture f;
  float one = 1.0f;
  auto impulse = c::Filter(8, &one, (&one) + 1);
  f.conv_input.add_block(test_signal);
  f.conv_output.set_filter(impulse);
  float* result = f.conv_output.convolve();
  CHECK_RANGE(result, test_signal, 8);
  f.conv_input.add_block(test_signal + 8);
  result = f.conv_output.convolve();
  CHECK_RANGE(result, test_signal + 8, 8);
}

static void section_and_more()
{
  Fixture f;
  f.conv_output.set_filter(f.filter);
  float input[8] = { 0.0f };
  input[1] = 1.0f;
  f.conv_input.add_block(input);
  float* result = f.conv_output.convolve();
  CHECK_RANGE(result, zeros, 8);

  input[1] = 2.0f;
  f.conv_input.add_block(input);
  CHECK(!f.conv_output.queues_empty());
  f.conv_output.rotate_queues();
  result = f.conv_output.convolve();
  float expected[8] = { 0.0f };
  expected[3] = 5.0f; expected[4] = 4.0f; expected[5] = 3.0f;
  CHECK_RANGE(result, expected, 8);

  input[1] = 0.0f;
  f.conv_input.add_block(input);
  CHECK(f.conv_output.queues_empty());
  result = f.conv_output.convolve();
  expected[3] = 10.0f; expected[4] = 8.0f; expected[5] = 6.0f;
  CHECK_RANGE(result, expected, 8);

  f.conv_input.add_block(input);
  CHECK(f.conv_output.queues_empty());
  result = f.conv_output.convolve();
  CHECK_RANGE(result, zeros, 8);
  CHECK(f.conv_output.queues_empty());
}

static void section_static_output_impulse()
{
  float one = 1.0f;
  auto sconv_input = c::Input(8, 1);
  auto sconv_output = c::StaticOutput(sconv_input, &one, (&one) + 1);
  sconv_input.add_block(test_signal);
  float* result = sconv_output.convolve();
  CHECK_RANGE(result, test_signal, 8);
  sconv_input.add_block(test_signal + 8);
  result = sconv_output.convolve();
  CHECK_RANGE(result, test_signal + 8, 8);
}

static void section_static_output_frequency_domain()
{
  Fixture f;
  float one = 1.0f;
  auto fd_filter = c::Filter(8, 3);     // 3 partitions (only 1 is really needed)
  auto sconv_input = c::Input(8, 7);    // 7 partitions (just for fun)
  f.conv_input.prepare_filter(&one, (&one) + 1, fd_filter);
  auto sconv_output = c::StaticOutput(sconv_input, fd_filter);
  CHECK(sconv_input.partitions() == 7);
  CHECK(fd_filter.partitions() == 3);
  sconv_input.add_block(test_signal);
  float* result = sconv_output.convolve();
  CHECK_RANGE(result, test_signal, 8);
}

static void section_combinations()
{
  Fixture f;
  auto so = c::StaticOutput(f.conv_input, f.filter);
  auto conv = c::Convolver(8, f.partitions);
  conv.set_filter(f.filter);
  conv.rotate_queues();
  auto sconv = c::StaticConvolver(f.filter, f.partitions);

  float input[8] = { 0.0f };
  float expected[4][8] = { {0}, {0, 0, 0, 5, 4, 3, 0, 0}, {0, 0, 0, 10, 8, 6, 0, 0}, {0} };
  float values[4] = { 1.0f, 2.0f, 0.0f, 0.0f };
  for (int step = 0; step < 4; ++step)
  {
    input[1] = values[step];
    f.conv_input.add_block(input);
    conv.add_block(input);
    sconv.add_block(input);
    float* result = so.convolve();
    CHECK_RANGE(result, expected[step], 8);
    result = conv.convolve();
    CHECK_RANGE(result, expected[step], 8);
    result = sconv.convolve();
    CHECK_RANGE(result, expected[step], 8);
  }
}

static void section_errors_and_extras()
{
  // "Convolver: block size must be a multiple of 8!" (convolver.h:163-166)
  bool thrown = false;
  try { c::Input bad(12, 1); }
  catch (const std::logic_error& e) { thrown = std::string(e.what()).find("multiple of 8") != std::string::npos; }
  CHECK(thrown);
  CHECK(c::min_partitions(1024, 48000) == 47);

  // weight argument and transform_nested(lerp)
  float a[8] = { 1, 0, 0, 0, 0, 0, 0, 0 };
  float b[8] = { 0, 1, 0, 0, 0, 0, 0, 0 };
  c::Filter fa(8, a, a + 8), fb(8, b, b + 8), mix(8, 1);
  c::transform_nested(fa, fb, mix, c::lerp(0.25f));
  c::Input in(8, 1);
  c::StaticOutput out(in, mix);
  float x[8] = { 1, 0, 0, 0, 0, 0, 0, 0 };
  in.add_block(x);
  float* y = out.convolve(2.0f);
  float exp[8] = { 1.5f, 0.5f, 0, 0, 0, 0, 0, 0 };
  CHECK_RANGE(y, exp, 8);

  // transform_nested with an arbitrary binary host callable (convolver.h:773-817):
  // element-wise on the spectra, zero-flag rules for unequal lengths.  in1 has two
  // partitions, in2 one, out three: partition 0 = f(a, b), partition 1 = f(a, 0),
  // partition 2 is zero-flagged.  With the linear f(u, v) = 2u - 3v the result is
  // the filter 2 * ha - 3 * hb.
  float ha[16], hb[8];
  for (int i = 0; i < 16; ++i) ha[i] = 0.01f * float(i + 1);
  for (int i = 0; i < 8; ++i) hb[i] = 0.05f * float(8 - i);
  c::Filter f1(8, ha, ha + 16), f2(8, hb, hb + 8), res(8, 3);
  c::transform_nested(f1, f2, res, [](float u, float v) { return 2.0f * u - 3.0f * v; });
  std::vector<float> spec(16);
  CHECK(!res.download(0, spec.data()));
  CHECK(!res.download(1, spec.data()));
  CHECK(res.download(2, spec.data()));   // both inputs past their end -> zero flag
  c::Input in2(8, 3);
  c::StaticOutput out2(in2, res);
  in2.add_block(x);                      // unit impulse -> the impulse response, block by block
  float* y0 = out2.convolve();
  float exp0[8], exp1[8];
  for (int i = 0; i < 8; ++i) { exp0[i] = 2.0f * ha[i] - 3.0f * hb[i]; exp1[i] = 2.0f * ha[8 + i]; }
  CHECK_RANGE(y0, exp0, 8);
  float zeros[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  in2.add_block(zeros);
  float* y1 = out2.convolve();
  CHECK_RANGE(y1, exp1, 8);
}

int main()
{
  try
  {
    section_silence();
    section_impulse();
    section_and_more();
    section_static_output_impulse();
    section_static_output_frequency_domain();
    section_combinations();
    section_errors_and_extras();
  }
  catch (const std::exception& e)
  {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("%s (%d assertions, %d failed)\n", failures ? "FAILED" : "All tests passed", checks, failures);
  return failures ? 1 : 0;
}
