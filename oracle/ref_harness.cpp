/* ref_harness.cpp -- C-ABI harness over the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  This file compiles the reference's own
 * apf/apf/convolver.h, apf/apf/combine_channels.h and
 * src/{generic,binaural,brs,wfs}renderer.h from where they lie under
 * /root/reference (see oracle/Makefile) into oracle/_ref/libssr_ref.so.
 * Nothing under the product package links or loads it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Absent third-party libraries are replaced by oracle/refshim/ (FFTW3 ->
 * fftw3.h shim with a hand-written float32 radix-4 Stockham FFT, libsndfile ->
 * in-memory registry, libxml2 -> typedef stubs).  No arithmetic line of the
 * reference is touched.
 *
 * Policies: apf::pointer_policy<float*> (apf/apf/pointer_policy.h:112-122; the
 * caller drives blocks) and apf::posix_thread_policy, selected the same way the
 * MEX wrapper does it (mex/ssr_nfc_hoa.cpp:33-39): first policy header wins.
 */

#include "apf/pointer_policy.h"
#include "apf/posix_thread_policy.h"

#include "apf/convolver.h"
#include "apf/combine_channels.h"

#include "genericrenderer.h"
#include "binauralrenderer.h"
#include "brsrenderer.h"
#include "ssr_global.h"    // ssr::c_inverse, which wfsrenderer.h uses without including it (the application gets it through src/controller.h:45)
#include "wfsrenderer.h"   // its pre-filter apf::conv::StaticConvolver per input (src/wfsrenderer.h:104-123)

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace
{
thread_local std::string g_error;

template<typename F>
int guarded(F f)
{
  try { f(); return 0; }
  catch (const std::exception& e) { g_error = e.what(); return -1; }
  catch (...) { g_error = "unknown exception"; return -1; }
}

namespace c = apf::conv;

// OutputBase has a protected ctor and convolve() is public; Output and
// StaticOutput are distinct leaf classes, so keep a tagged holder.
struct OutHolder
{
  std::unique_ptr<c::Output> dyn;
  std::unique_ptr<c::StaticOutput> stat;
  c::OutputBase* base() { return dyn ? static_cast<c::OutputBase*>(dyn.get())
                                     : static_cast<c::OutputBase*>(stat.get()); }
};

struct IRenderer
{
  virtual ~IRenderer() {}
  virtual void load_reproduction_setup() = 0;
  virtual void add_output() = 0;
  virtual void add_loudspeaker(float x, float y, float azimuth, int subwoofer) = 0;
  virtual int add_source(const std::string& file) = 0;
  virtual void rem_source(int id) = 0;
  virtual void activate() = 0;
  virtual int n_in() const = 0;
  virtual int n_out() const = 0;
  virtual int block_size() const = 0;
  virtual void callback(float* const* in, float* const* out) = 0;
  virtual void set_source(int id, const char* field, float a, float b) = 0;
  virtual void set_state(const char* field, float a, float b) = 0;
};

template<typename R> struct add_output_helper
{
  static void add(R& r) { r.add(typename R::Output::Params()); }
};

// Loudspeaker-based renderers (src/loudspeakerrenderer.h:49-58): an output is a Loudspeaker with
// position, orientation and model; the XML loader is replaced by explicit calls.
template<typename R> struct loudspeaker_helper
{
  static void add(R&, float, float, float, int)
  { throw std::logic_error("ref_harness: this renderer has no loudspeakers"); }
};
template<> struct loudspeaker_helper<ssr::WfsRenderer>
{
  static void add(ssr::WfsRenderer& r, float x, float y, float azimuth, int subwoofer)
  {
    ssr::WfsRenderer::Output::Params p;
    p.position = Position(x, y);
    p.orientation = Orientation(azimuth);
    p.model = subwoofer ? Loudspeaker::subwoofer : Loudspeaker::normal;
    r.add(p);
  }
};

// GenericRenderer::load_reproduction_setup() is LoudspeakerRenderer's XML loader
// (src/loudspeakerrenderer.h:95-150); libxml2 is absent, so it must never be
// instantiated.  Outputs are added with add_output() instead.
template<typename R> struct setup_helper
{
  static void load(R& r) { r.load_reproduction_setup(); }
};
template<> struct setup_helper<ssr::GenericRenderer>
{
  static void load(ssr::GenericRenderer&)
  {
    throw std::logic_error("ref_harness: GenericRenderer reproduction setup "
        "needs libxml2; use ref_renderer_add_outputs()");
  }
};

template<> struct setup_helper<ssr::WfsRenderer>
{
  static void load(ssr::WfsRenderer&)
  {
    throw std::logic_error("ref_harness: WfsRenderer reproduction setup "
        "needs libxml2; use ref_renderer_add_loudspeaker()");
  }
};

template<typename R>
struct RendererWrap : IRenderer
{
  explicit RendererWrap(const apf::parameter_map& p) : r(p) {}

  void load_reproduction_setup() override { setup_helper<R>::load(r); }
  void add_output() override { add_output_helper<R>::add(r); }
  void add_loudspeaker(float x, float y, float azimuth, int subwoofer) override
  { loudspeaker_helper<R>::add(r, x, y, azimuth, subwoofer); }

  int add_source(const std::string& file) override
  {
    apf::parameter_map p;
    if (!file.empty()) p.set("properties_file", file);
    return r.add_source(p);
  }

  void rem_source(int id) override
  {
    auto* s = r.get_source(id);
    if (!s) throw std::logic_error("ref_harness: no such source");
    r.rem_source(s);
  }

  void activate() override { r.activate(); }
  int n_in() const override { return r.in_channels(); }
  int n_out() const override { return r.out_channels(); }
  int block_size() const override { return r.block_size(); }

  void callback(float* const* in, float* const* out) override
  {
    r.audio_callback(r.block_size(), in, out);
  }

  void set_source(int id, const char* field, float a, float b) override
  {
    auto* s = r.get_source(id);
    if (!s) throw std::logic_error("ref_harness: no such source");
    std::string f(field);
    // same locking discipline as src/rendersubscriber.h:65-212
    auto guard = r.get_scoped_lock();
    if (f == "gain") s->gain = a;
    else if (f == "mute") s->mute = (a != 0.0f);
    else if (f == "position") s->position = Position(a, b);
    else if (f == "orientation") s->orientation = Orientation(a);
    else if (f == "model") s->model = (a != 0.0f) ? ::Source::plane
                                                   : ::Source::point;
    else throw std::logic_error("ref_harness: unknown source field " + f);
  }

  void set_state(const char* field, float a, float b) override
  {
    std::string f(field);
    auto guard = r.get_scoped_lock();
    if (f == "reference_position") r.state.reference_position = Position(a, b);
    else if (f == "reference_orientation")
      r.state.reference_orientation = Orientation(a);
    else if (f == "reference_offset_position")
      r.state.reference_offset_position = Position(a, b);
    else if (f == "reference_offset_orientation")
      r.state.reference_offset_orientation = Orientation(a);
    else if (f == "master_volume") r.state.master_volume = a;
    else if (f == "processing") r.state.processing = (a != 0.0f);
    else if (f == "amplitude_reference_distance")
      r.state.amplitude_reference_distance = a;
    else throw std::logic_error("ref_harness: unknown state field " + f);
  }

  R r;
};

struct RendererHandle
{
  std::unique_ptr<IRenderer> impl;
  std::vector<float*> in_ptrs, out_ptrs;
};

}  // namespace

extern "C"
{

const char* ref_last_error() { return g_error.c_str(); }

const char* ref_describe()
{
  return "reference apf::conv + ssr renderers, unmodified headers; "
         "FFTW3 replaced by shim FFT (hand-written float32 radix-4 Stockham)";
}

/* ---------- in-memory sound files (replaces libsndfile) ---------- */

int ref_sndfile_register(const char* name, const float* interleaved
    , long frames, int channels, int samplerate, int copy)
{
  return guarded([&] {
    auto f = std::make_shared<ssr_b200_sndshim::MemFile>();
    f->channels = channels; f->samplerate = samplerate; f->frames = frames;
    if (copy)
    {
      f->owned.assign(interleaved, interleaved + size_t(frames) * channels);
      f->data = f->owned.data();
    }
    else f->data = interleaved;
    ssr_b200_sndshim::registry()[name] = f;
  });
}

void ref_sndfile_unregister(const char* name)
{
  ssr_b200_sndshim::registry().erase(name);
}

/* ---------- Level 1: apf::conv objects ---------- */

long ref_min_partitions(long block, long taps)
{
  return long(c::min_partitions(size_t(block), size_t(taps)));
}

void* ref_filter_new_time(int block, const float* h, long n, int partitions)
{
  void* out = nullptr;
  guarded([&] { out = new c::Filter(size_t(block), h, h + n
        , size_t(partitions)); });
  return out;
}

void* ref_filter_new_empty(int block, int partitions)
{
  void* out = nullptr;
  guarded([&] { out = new c::Filter(size_t(block), size_t(partitions)); });
  return out;
}

int ref_filter_partitions(void* f)
{ return int(static_cast<c::Filter*>(f)->partitions()); }

/* copies the sorted spectrum of partition p; returns the zero flag */
int ref_filter_get(void* f, int p, float* out)
{
#ifdef SSR_REF_GPU_FACADE
  // built against ssr-pre-0.4.0_b200/cpp/apf/convolver.h (the GPU engine's facade of
  // the same class API): spectra live in HBM, fetched in the reference's layout
  return static_cast<c::Filter*>(f)->download(size_t(p), out) ? 1 : 0;
#else
  auto& node = (*static_cast<c::Filter*>(f))[size_t(p)];
  std::copy(node.begin(), node.end(), out);
  return node.zero ? 1 : 0;
#endif
}

void ref_filter_delete(void* f) { delete static_cast<c::Filter*>(f); }

void* ref_transform_new(int block)
{
  void* out = nullptr;
  guarded([&] { out = new c::Transform(size_t(block)); });
  return out;
}

void ref_transform_prepare_filter(void* t, const float* h, long n, void* f)
{
  static_cast<c::Transform*>(t)->prepare_filter(h, h + n
      , *static_cast<c::Filter*>(f));
}

void ref_transform_delete(void* t) { delete static_cast<c::Transform*>(t); }

void* ref_input_new(int block, int partitions)
{
  void* out = nullptr;
  guarded([&] { out = new c::Input(size_t(block), size_t(partitions)); });
  return out;
}

void ref_input_add_block(void* in, const float* x)
{ static_cast<c::Input*>(in)->add_block(x); }

int ref_input_partitions(void* in)
{ return int(static_cast<c::Input*>(in)->partitions()); }

/* spectrum p of the FDL (0 = newest); returns zero flag */
int ref_input_get(void* in, int p, float* out)
{
#ifdef SSR_REF_GPU_FACADE
  if (ssr_input_download(static_cast<c::Input*>(in)->handle(), p, out) != SSR_OK) return -1;
  return 0;   // the engine keeps no zero flag for delay-line spectra (an all-zero block is stored as zeros)
#else
  auto it = static_cast<c::Input*>(in)->spectra.begin();
  std::advance(it, p);
  std::copy(it->begin(), it->end(), out);
  return it->zero ? 1 : 0;
#endif
}

void ref_input_delete(void* in) { delete static_cast<c::Input*>(in); }

void* ref_output_new(void* in)
{
  auto* h = new OutHolder;
  if (guarded([&] { h->dyn.reset(new c::Output(*static_cast<c::Input*>(in))); }))
  { delete h; return nullptr; }
  return h;
}

void* ref_static_output_new_time(void* in, const float* h, long n)
{
  auto* o = new OutHolder;
  if (guarded([&] { o->stat.reset(new c::StaticOutput(
            *static_cast<c::Input*>(in), h, h + n)); }))
  { delete o; return nullptr; }
  return o;
}

void* ref_static_output_new_filter(void* in, void* f)
{
  auto* o = new OutHolder;
  if (guarded([&] { o->stat.reset(new c::StaticOutput(
            *static_cast<c::Input*>(in), *static_cast<c::Filter*>(f))); }))
  { delete o; return nullptr; }
  return o;
}

void ref_output_set_filter(void* o, void* f)
{ static_cast<OutHolder*>(o)->dyn->set_filter(*static_cast<c::Filter*>(f)); }

int ref_output_queues_empty(void* o)
{ return static_cast<OutHolder*>(o)->dyn->queues_empty() ? 1 : 0; }

void ref_output_rotate_queues(void* o)
{ static_cast<OutHolder*>(o)->dyn->rotate_queues(); }

void ref_output_convolve(void* o, float weight, float* y)
{
  auto* base = static_cast<OutHolder*>(o)->base();
  float* r = base->convolve(weight);
  std::copy(r, r + base->block_size(), y);
}

void ref_output_delete(void* o) { delete static_cast<OutHolder*>(o); }

/* transform_nested with the lerp of src/binauralrenderer.h:372-377 */
void ref_transform_nested_lerp(void* a, void* b, void* out, float t)
{
  c::transform_nested(*static_cast<c::Filter*>(a), *static_cast<c::Filter*>(b)
      , *static_cast<c::Filter*>(out)
      , [t] (float one, float two) { return (1.0f - t) * one + t * two; });
}

/* apf::raised_cosine_fade (apf/apf/combine_channels.h:470-497) */
void ref_fade_table(int block, float* fade_out, float* fade_in)
{
  apf::raised_cosine_fade<float> fade{size_t(block)};
  std::copy(fade.fade_out_begin(), fade.fade_out_begin() + block, fade_out);
  std::copy(fade.fade_in_begin(), fade.fade_in_begin() + block, fade_in);
}

/* Whole-signal StaticConvolver run (config 1 style); returns seconds spent in
 * the add_block/convolve loop (steady_clock wall time). */
double ref_static_convolver_run(int block, const float* h, long taps
    , const float* x, long blocks, float* y, int repeat)
{
  double secs = -1.0;
  guarded([&] {
    c::StaticConvolver conv(size_t(block), h, h + taps);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < (repeat > 0 ? repeat : 1); ++r)
    {
      for (long b = 0; b < blocks; ++b)
      {
        conv.add_block(x + b * block);
        const float* res = conv.convolve();
        if (y) std::copy(res, res + block, y + b * block);
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    secs = std::chrono::duration<double>(t1 - t0).count();
  });
  return secs;
}

/* ---------- Level 2: renderers under pointer_policy ---------- */

void* ref_renderer_new(const char* kind, int block, int sample_rate
    , int threads, const char* hrir_file, long hrir_size)
{
  auto* h = new RendererHandle;
  int rc = guarded([&] {
    apf::parameter_map p;
    p.set("block_size", block);
    p.set("sample_rate", sample_rate);
    p.set("threads", threads);
    if (hrir_file && hrir_file[0]) p.set("hrir_file", std::string(hrir_file));
    if (hrir_size > 0) p.set("hrir_size", hrir_size);
    std::string k(kind);
    if (k == "generic") h->impl.reset(new RendererWrap<ssr::GenericRenderer>(p));
    else if (k == "binaural")
      h->impl.reset(new RendererWrap<ssr::BinauralRenderer>(p));
    else if (k == "brs") h->impl.reset(new RendererWrap<ssr::BrsRenderer>(p));
    else throw std::logic_error("ref_harness: unknown renderer kind " + k);
  });
  if (rc) { delete h; return nullptr; }
  return h;
}

/* WfsRenderer (src/wfsrenderer.h:60-88): pre-filter from `prefilter_file` (one channel), the
 * delay line's size and initial delay in samples (src/configuration.h:62-66). */
void* ref_renderer_new_wfs(int block, int sample_rate, int threads, const char* prefilter_file
    , long delayline_size, long initial_delay)
{
  auto* h = new RendererHandle;
  int rc = guarded([&] {
    apf::parameter_map p;
    p.set("block_size", block);
    p.set("sample_rate", sample_rate);
    p.set("threads", threads);
    p.set("prefilter_file", std::string(prefilter_file));
    p.set("delayline_size", delayline_size);
    p.set("initial_delay", initial_delay);
    h->impl.reset(new RendererWrap<ssr::WfsRenderer>(p));
  });
  if (rc) { delete h; return nullptr; }
  return h;
}

int ref_renderer_add_loudspeaker(void* r, float x, float y, float azimuth, int subwoofer)
{
  return guarded([&] {
      static_cast<RendererHandle*>(r)->impl->add_loudspeaker(x, y, azimuth, subwoofer); });
}

void ref_renderer_delete(void* r) { delete static_cast<RendererHandle*>(r); }

int ref_renderer_load_reproduction_setup(void* r)
{
  return guarded([&] {
      static_cast<RendererHandle*>(r)->impl->load_reproduction_setup(); });
}

int ref_renderer_add_outputs(void* r, int m)
{
  return guarded([&] {
      for (int i = 0; i < m; ++i)
        static_cast<RendererHandle*>(r)->impl->add_output(); });
}

int ref_renderer_add_source(void* r, const char* properties_file)
{
  int id = -1;
  guarded([&] { id = static_cast<RendererHandle*>(r)->impl->add_source(
        properties_file ? properties_file : ""); });
  return id;
}

int ref_renderer_rem_source(void* r, int id)
{
  return guarded([&] { static_cast<RendererHandle*>(r)->impl->rem_source(id); });
}

int ref_renderer_activate(void* r)
{
  return guarded([&] { static_cast<RendererHandle*>(r)->impl->activate(); });
}

int ref_renderer_in_channels(void* r)
{ return static_cast<RendererHandle*>(r)->impl->n_in(); }

int ref_renderer_out_channels(void* r)
{ return static_cast<RendererHandle*>(r)->impl->n_out(); }

int ref_renderer_set_source(void* r, int id, const char* field, float a, float b)
{
  return guarded([&] {
      static_cast<RendererHandle*>(r)->impl->set_source(id, field, a, b); });
}

int ref_renderer_set_state(void* r, const char* field, float a, float b)
{
  return guarded([&] {
      static_cast<RendererHandle*>(r)->impl->set_state(field, a, b); });
}

/* One audio block: in = n_in x B contiguous, out = n_out x B contiguous. */
int ref_renderer_process(void* r, const float* in, float* out)
{
  auto* h = static_cast<RendererHandle*>(r);
  return guarded([&] {
    const int B = h->impl->block_size();
    const int ni = h->impl->n_in(), no = h->impl->n_out();
    h->in_ptrs.resize(ni); h->out_ptrs.resize(no);
    for (int i = 0; i < ni; ++i) h->in_ptrs[i] = const_cast<float*>(in) + size_t(i) * B;
    for (int i = 0; i < no; ++i) h->out_ptrs[i] = out + size_t(i) * B;
    h->impl->callback(h->in_ptrs.data(), h->out_ptrs.data());
  });
}

/* Timed block loop.  in = in_blocks x n_in x B (cycled), out = blocks x n_out x B
 * or NULL (then a single scratch block is overwritten).  azimuth (optional,
 * length `blocks`) sets state.reference_orientation before each block (head
 * rotation).  per_block_ms (optional, length `blocks`) receives the
 * wall time of each callback.  Returns total wall seconds (steady_clock). */
double ref_renderer_run(void* r, const float* in, long in_blocks, float* out
    , long blocks, const float* azimuth, double* per_block_ms)
{
  auto* h = static_cast<RendererHandle*>(r);
  double secs = -1.0;
  guarded([&] {
    const int B = h->impl->block_size();
    const int ni = h->impl->n_in(), no = h->impl->n_out();
    std::vector<float> scratch(out ? 0 : size_t(no) * B);
    h->in_ptrs.resize(ni); h->out_ptrs.resize(no);
    auto t0 = std::chrono::steady_clock::now();
    for (long b = 0; b < blocks; ++b)
    {
      auto tb = std::chrono::steady_clock::now();
      const float* ib = in + size_t(b % in_blocks) * ni * B;
      float* ob = out ? out + size_t(b) * no * B : scratch.data();
      for (int i = 0; i < ni; ++i) h->in_ptrs[i] = const_cast<float*>(ib) + size_t(i) * B;
      for (int i = 0; i < no; ++i) h->out_ptrs[i] = ob + size_t(i) * B;
      if (azimuth) h->impl->set_state("reference_orientation", azimuth[b], 0.0f);
      h->impl->callback(h->in_ptrs.data(), h->out_ptrs.data());
      if (per_block_ms)
      {
        auto te = std::chrono::steady_clock::now();
        per_block_ms[b] = std::chrono::duration<double, std::milli>(te - tb).count();
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    secs = std::chrono::duration<double>(t1 - t0).count();
  });
  return secs;
}

}  // extern "C"
