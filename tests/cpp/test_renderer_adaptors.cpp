// C++ test of the renderer drop-in adaptors (ssr-pre-0.4.0_b200/cpp/ssr/renderers.h)
// on a real GPU: the reference's usage pattern under apf::pointer_policy
// (parameter_map -> renderer -> add_source("properties_file") -> audio_callback)
// checked against float64 direct convolution computed here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ssr/renderers.h"

static int failures = 0, checks = 0;
#define CHECK(cond) do { ++checks; if (!(cond)) { ++failures; \
  std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)

static unsigned seed_ = 777u;
static float rnd() { seed_ = seed_ * 1664525u + 1013904223u; return float(seed_ >> 8) / 8388608.0f - 1.0f; }

static void put32(std::FILE* f, uint32_t v) { std::fwrite(&v, 4, 1, f); }
static void put16(std::FILE* f, uint16_t v) { std::fwrite(&v, 2, 1, f); }

// frames x channels interleaved
static void write_wav(const std::string& path, const std::vector<float>& data, int channels, int fs, bool pcm16)
{
  std::FILE* f = std::fopen(path.c_str(), "wb");
  const uint32_t bps = pcm16 ? 2 : 4;
  const uint32_t bytes = uint32_t(data.size()) * bps;
  std::fwrite("RIFF", 1, 4, f); put32(f, 36 + bytes); std::fwrite("WAVE", 1, 4, f);
  std::fwrite("fmt ", 1, 4, f); put32(f, 16); put16(f, pcm16 ? 1 : 3); put16(f, uint16_t(channels));
  put32(f, uint32_t(fs)); put32(f, uint32_t(fs) * channels * bps); put16(f, uint16_t(channels * bps)); put16(f, uint16_t(8 * bps));
  std::fwrite("data", 1, 4, f); put32(f, bytes);
  for (float v : data)
  {
    if (pcm16) { int16_t s = int16_t(std::lrint(v * 32768.0f)); std::fwrite(&s, 2, 1, f); }
    else std::fwrite(&v, 4, 1, f);
  }
  std::fclose(f);
}

static std::vector<double> direct(const std::vector<float>& x, const std::vector<float>& h)
{
  std::vector<double> y(x.size(), 0.0);
  for (size_t i = 0; i < x.size(); ++i)
    for (size_t k = 0; k < h.size() && k <= i; ++k) y[i] += double(x[i - k]) * double(h[k]);
  return y;
}

static std::vector<float> column(const std::vector<float>& inter, int channels, int c)
{
  std::vector<float> out(inter.size() / channels);
  for (size_t t = 0; t < out.size(); ++t) out[t] = inter[t * channels + c];
  return out;
}

static const int B = 64, FS = 48000, T = 6;

static apf::parameter_map base_params()
{
  apf::parameter_map p;
  p.set("block_size", B);
  p.set("sample_rate", FS);
  p.set("threads", 4);  // accepted, ignored
  return p;
}

static void test_generic(const std::string& dir)
{
  const int N = 2, M = 3, L = 100;
  std::vector<std::vector<float>> irs(N);
  auto p = base_params();
  p.set("outputs", M);
  ssr::GenericRenderer r(p);
  CHECK(std::string(r.name()) == "GenericRenderer");
  for (int n = 0; n < N; ++n)
  {
    irs[n].resize(size_t(L) * M);
    // multiples of 2^-15 so that the PCM16 file (source 1) is exact
    for (auto& v : irs[n]) v = std::floor(rnd() * 0.1f * 32768.0f) / 32768.0f;
    const std::string path = dir + "/generic_ir_" + std::to_string(n) + ".wav";
    write_wav(path, irs[n], M, FS, n == 1);
    apf::parameter_map sp;
    sp.set("properties_file", path);
    const int id = r.add_source(sp);
    CHECK(r.get_source(id) != nullptr);
  }
  CHECK(r.in_channels() == N && r.out_channels() == M);
  r.activate();
  std::vector<std::vector<float>> x(N, std::vector<float>(size_t(T) * B));
  for (auto& s : x) for (auto& v : s) v = rnd();
  std::vector<std::vector<float>> y(M, std::vector<float>(size_t(T) * B));
  for (int t = 0; t < T; ++t)
  {
    float* in[N]; float* out[M];
    for (int n = 0; n < N; ++n) in[n] = x[n].data() + size_t(t) * B;
    for (int m = 0; m < M; ++m) out[m] = y[m].data() + size_t(t) * B;
    r.audio_callback(B, in, out);
  }
  double worst = 0, worst0 = 0;
  for (int m = 0; m < M; ++m)
  {
    std::vector<double> truth(size_t(T) * B, 0.0);
    for (int n = 0; n < N; ++n)
    {
      auto d = direct(x[n], column(irs[n], M, m));
      for (size_t i = 0; i < truth.size(); ++i) truth[i] += d[i];
    }
    for (size_t i = B; i < truth.size(); ++i) worst = std::fmax(worst, std::fabs(truth[i] - y[m][i]));
    // first block: fade-in from weight 0 (src/rendererbase.h:379, raised cosine)
    for (int i = 0; i < B; ++i)
    {
      const double fi = 0.5 * std::cos(M_PI * double(B - i) / B) + 0.5;
      worst0 = std::fmax(worst0, std::fabs(truth[i] * fi - y[m][i]));
    }
  }
  std::printf("generic: max err %.3g (blocks >= 1), %.3g (fade-in block)\n", worst, worst0);
  CHECK(worst <= 1e-5);
  CHECK(worst0 <= 1e-5);

  // control change: mute -> crossfade to silence, then zeros
  r.get_source(1)->mute = true;
  r.get_source(2)->mute = true;
  float* in[N]; float* out[M];
  std::vector<float> o(size_t(M) * B);
  for (int n = 0; n < N; ++n) in[n] = x[n].data();
  for (int m = 0; m < M; ++m) out[m] = o.data() + size_t(m) * B;
  r.audio_callback(B, in, out);  // fade-out block
  r.audio_callback(B, in, out);
  bool all_zero = true;
  for (float v : o) all_zero = all_zero && v == 0.0f;
  CHECK(all_zero);

  // error behaviour of load_sndfile / the Source constructor
  bool ok = false;
  try { apf::parameter_map sp; sp.set("properties_file", dir + "/does_not_exist.wav"); r.add_source(sp); }
  catch (const std::logic_error& e) { ok = std::string(e.what()).find("couldn't be loaded") != std::string::npos; }
  CHECK(ok);
  ok = false;
  try
  {
    write_wav(dir + "/two_ch.wav", std::vector<float>(20, 0.1f), 2, FS, false);
    apf::parameter_map sp; sp.set("properties_file", dir + "/two_ch.wav"); r.add_source(sp);
  }
  catch (const std::logic_error& e) { ok = std::string(e.what()).find("channels instead of") != std::string::npos; }
  CHECK(ok);
  ok = false;
  try
  {
    write_wav(dir + "/wrong_fs.wav", std::vector<float>(30, 0.1f), 3, 44100, false);
    apf::parameter_map sp; sp.set("properties_file", dir + "/wrong_fs.wav"); r.add_source(sp);
  }
  catch (const std::logic_error& e) { ok = std::string(e.what()).find("has sample rate 44100 instead of 48000") != std::string::npos; }
  CHECK(ok);
  ok = false;
  try { r.add_source(); }
  catch (const std::logic_error& e) { ok = std::string(e.what()).find("Empty file name") != std::string::npos; }
  CHECK(ok);
}

static void test_brs_and_binaural(const std::string& dir)
{
  const int A = 4, L = 150;
  std::vector<float> set(size_t(L) * 2 * A);
  for (auto& v : set) v = rnd() * 0.1f;
  const std::string path = dir + "/set.wav";
  write_wav(path, set, 2 * A, FS, false);
  std::vector<float> x(size_t(T) * B);
  for (auto& v : x) v = rnd();

  {
    ssr::BrsRenderer r(base_params());
    r.load_reproduction_setup();
    apf::parameter_map sp; sp.set("properties_file", path);
    r.add_source(sp);
    r.activate();
    // 90 degrees is the middle of index 0 (src/brsrenderer.h:152-155)
    r.state.reference_orientation = Orientation(90.0f);
    std::vector<float> y0(size_t(T) * B), y1(size_t(T) * B);
    for (int t = 0; t < T; ++t)
    {
      float* in[1] = { x.data() + size_t(t) * B };
      float* out[2] = { y0.data() + size_t(t) * B, y1.data() + size_t(t) * B };
      r.audio_callback(B, in, out);
    }
    // after the staggered switch has settled (P - 1 = 2 rotations) the output is
    // the plain convolution with BRIR pair 0
    auto t0 = direct(x, column(set, 2 * A, 0)), t1 = direct(x, column(set, 2 * A, 1));
    double worst = 0;
    for (size_t i = 3 * B; i < t0.size(); ++i)
      worst = std::fmax(worst, std::fmax(std::fabs(t0[i] - y0[i]), std::fabs(t1[i] - y1[i])));
    std::printf("brs: max err %.3g\n", worst);
    CHECK(worst <= 1e-5);
  }
  {
    auto p = base_params();
    p.set("hrir_file", path);
    p.set("hrir_size", 60);  // first 60 taps only -> P = 1
    ssr::BinauralRenderer r(p);
    r.load_reproduction_setup();
    const int id = r.add_source();
    r.activate();
    r.get_source(id)->position = Position(0.0f, 2.0f);  // 90 deg, 2 m -> index 1, weight 0.25
    std::vector<float> y0(size_t(T) * B), y1(size_t(T) * B);
    for (int t = 0; t < T; ++t)
    {
      float* in[1] = { x.data() + size_t(t) * B };
      float* out[2] = { y0.data() + size_t(t) * B, y1.data() + size_t(t) * B };
      r.audio_callback(B, in, out);
    }
    auto h0 = column(set, 2 * A, 2), h1 = column(set, 2 * A, 3);
    h0.resize(60); h1.resize(60);
    auto t0 = direct(x, h0), t1 = direct(x, h1);
    double worst = 0;
    for (size_t i = B; i < t0.size(); ++i)
      worst = std::fmax(worst, std::fmax(std::fabs(0.25 * t0[i] - y0[i]), std::fabs(0.25 * t1[i] - y1[i])));
    std::printf("binaural: max err %.3g\n", worst);
    CHECK(worst <= 1e-5);
  }
  // odd channel count (src/brsrenderer.h:101-105)
  bool ok = false;
  try
  {
    write_wav(dir + "/odd.wav", std::vector<float>(30, 0.1f), 3, FS, false);
    ssr::BrsRenderer r(base_params());
    apf::parameter_map sp; sp.set("properties_file", dir + "/odd.wav"); r.add_source(sp);
  }
  catch (const std::logic_error& e) { ok = std::string(e.what()).find("multiple of 2") != std::string::npos; }
  CHECK(ok);
}

// "devices" = several CUDA ordinals: the GenericRenderer shards its outputs over
// one engine per entry (here all on device 0 unless SSR_TEST_DEVICES lists others)
// and audio_callback() still delivers every output -- the single-process form of
// the multi-GPU path (the reference spreads its Output items over worker threads,
// apf/apf/mimoprocessor.h:452-483).
static void test_generic_sharded(const std::string& dir)
{
  const int N = 3, M = 7, L = 150, TT = 8;
  const char* env = std::getenv("SSR_TEST_DEVICES");
  for (const std::string& devices : {std::string(env ? env : "0,0"), std::string("0,0,0")})
  {
    auto p = base_params();
    p.set("outputs", M);
    ssr::GenericRenderer whole(p);
    p.set("devices", devices);
    ssr::GenericRenderer sharded(p);
    CHECK(sharded.shards() == int((devices.size() + 1) / 2));
    CHECK(sharded.out_channels() == M && whole.out_channels() == M);
    std::vector<std::vector<float>> irs(N);
    for (int n = 0; n < N; ++n)
    {
      irs[n].resize(size_t(L) * M);
      for (auto& v : irs[n]) v = rnd() * 0.1f;
      const std::string path = dir + "/sharded_ir_" + std::to_string(n) + ".wav";
      write_wav(path, irs[n], M, FS, false);
      apf::parameter_map sp;
      sp.set("properties_file", path);
      const int a = whole.add_source(sp), b = sharded.add_source(sp);
      CHECK(a == b);
    }
    CHECK(sharded.in_channels() == N);
    std::vector<std::vector<float>> x(N, std::vector<float>(size_t(TT) * B));
    for (auto& s : x) for (auto& v : s) v = rnd();
    double worst = 0, worst_truth = 0;
    std::vector<std::vector<float>> ys(M, std::vector<float>(size_t(TT) * B)), yw = ys;
    for (int t = 0; t < TT; ++t)
    {
      if (t == 3) { whole.get_source(2)->gain = 0.25f; sharded.get_source(2)->gain = 0.25f; }      // every shard
      if (t == 5) { whole.state.master_volume = 0.5f; sharded.state.master_volume = 0.5f; }
      if (t == 6) { whole.rem_source(whole.get_source(1)); sharded.rem_source(sharded.get_source(1)); }
      const int n_now = t >= 6 ? N - 1 : N;
      float* in[N]; float* o1[M]; float* o2[M];
      for (int n = 0; n < n_now; ++n) in[n] = x[t >= 6 ? n + 1 : n].data() + size_t(t) * B;
      for (int m = 0; m < M; ++m) { o1[m] = yw[m].data() + size_t(t) * B; o2[m] = ys[m].data() + size_t(t) * B; }
      whole.audio_callback(B, in, o1);
      sharded.audio_callback(B, in, o2);
      for (int m = 0; m < M; ++m)
        for (int i = 0; i < B; ++i) worst = std::fmax(worst, std::fabs(double(o1[m][i]) - double(o2[m][i])));
    }
    // blocks 1-2 (constant gains, all sources): against float64 direct convolution
    for (int m = 0; m < M; ++m)
    {
      std::vector<double> truth(size_t(TT) * B, 0.0);
      for (int n = 0; n < N; ++n)
      {
        auto d = direct(x[n], column(irs[n], M, m));
        for (size_t i = 0; i < truth.size(); ++i) truth[i] += d[i];
      }
      for (size_t i = B; i < size_t(3) * B; ++i) worst_truth = std::fmax(worst_truth, std::fabs(truth[i] - ys[m][i]));
    }
    std::printf("generic sharded over devices {%s}: max |sharded - whole| %.3g, vs direct convolution %.3g\n",
        devices.c_str(), worst, worst_truth);
    CHECK(worst <= 2e-6);
    CHECK(worst_truth <= 1e-5);
  }
  bool ok = false;
  try { auto p = base_params(); p.set("devices", "0,0"); p.set("hrir_file", "x.wav"); ssr::BrsRenderer r(p); }
  catch (const std::logic_error&) { ok = true; }
  CHECK(ok);
}

int main(int argc, char* argv[])
{
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  try
  {
    test_generic(dir);
    test_generic_sharded(dir);
    test_brs_and_binaural(dir);
  }
  catch (const std::exception& e)
  {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("%s (%d assertions, %d failed)\n", failures ? "FAILED" : "All tests passed", checks, failures);
  return failures ? 1 : 0;
}
