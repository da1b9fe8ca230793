// capi.cu -- the extern "C" boundary declared in include/ssr_b200.h.
// Nothing throws across it: every entry point maps exceptions to ssr_status and
// stores the message for ssr_last_error().
#include "engine.hpp"

#include <algorithm>
#include <new>

using namespace ssr_b200;

struct ssr_engine { Engine impl; explicit ssr_engine(const ssr_engine_config& c) : impl(c) {} };
struct ssr_filterset { FilterSet impl; ssr_filterset(Engine* e, int c, int p) : impl(e, c, p) {} };
struct ssr_input { Input impl; ssr_input(Engine* e, int p) : impl(e, p) {} };
struct ssr_output { Output impl; ssr_output(Input* i, ssr_output_kind k) : impl(i, k) {} };
struct ssr_gather { Gather impl; ssr_gather(Engine* e, int w, int r, int root, const int* c) : impl(e, w, r, root, c) {} };
struct ssr_hostgather { HostGather impl; ssr_hostgather(Engine* e, int w, int r, int root, const int* c, const char* n) : impl(e, w, r, root, c, n) {} };
struct ssr_renderer { Renderer impl; ssr_renderer(Engine* e, const ssr_renderer_config& c) : impl(e, c) {} };

namespace
{
thread_local std::string g_error;

template<typename F>
int guarded(F f)
{
  try { f(); return SSR_OK; }
  catch (const Error& e) { g_error = e.what(); return e.code; }
  catch (const std::bad_alloc&) { g_error = "out of host memory"; return SSR_ERR_NOMEM; }
  catch (const std::exception& e) { g_error = e.what(); return SSR_ERR_INVALID; }
  catch (...) { g_error = "unknown error"; return SSR_ERR_INVALID; }
}

void require(bool ok, const char* what)
{
  if (!ok) throw Error(SSR_ERR_INVALID, std::string("ssr_b200: ") + what);
}

long min_partitions(long B, long taps) { return (taps + B - 1) / B; }  // convolver.h:59-62
}  // namespace

extern "C"
{

const char* ssr_last_error(void) { return g_error.c_str(); }
int ssr_abi_version(void) { return SSR_B200_ABI_VERSION; }

/* ------------------------------------------------------------------ engine */

int ssr_engine_create(const ssr_engine_config* cfg, ssr_engine** out)
{
  return guarded([&] {
    require(cfg && out, "engine_create: null argument");
    *out = nullptr;
    *out = new ssr_engine(*cfg);
  });
}

void ssr_engine_destroy(ssr_engine* e) { delete e; }
int ssr_engine_block_size(const ssr_engine* e) { return e ? e->impl.B : 0; }

int ssr_engine_synchronize(ssr_engine* e)
{
  return guarded([&] { require(e, "null engine"); e->impl.use_device(); e->impl.sync(); });
}

int ssr_engine_set_timing(ssr_engine* e, int enabled)
{
  return guarded([&] {
    require(e, "null engine");
    e->impl.use_device();
    e->impl.fold_stats();
    e->impl.timing = enabled != 0;
    e->impl.timing_every = enabled > 1 ? enabled : 1;
    e->impl.timing_counter = 0;
  });
}

int ssr_engine_get_stats(ssr_engine* e, ssr_engine_stats* out, int reset)
{
  return guarded([&] {
    require(e && out, "get_stats: null argument");
    e->impl.use_device();
    e->impl.fold_stats();
    *out = e->impl.stats;
    out->timing_enabled = e->impl.timing ? 1 : 0;
    out->timing_every = e->impl.timing_every;
    if (reset) { const int grid = e->impl.stats.last_mac_grid; e->impl.stats = ssr_engine_stats{}; e->impl.stats.last_mac_grid = grid; }
  });
}

int ssr_engine_debug_mac_trace(ssr_engine* e, uint64_t* out, int max_ctas)
{
  return guarded([&] {
    require(e && out && max_ctas > 0, "debug_mac_trace: bad argument");
    Engine& E = e->impl;
    if (!E.mac_trace.ptr) throw Error(SSR_ERR_STATE, "ssr_b200: this engine does not run the bulk-copy MAC (B < 1024)");
    E.use_device();
    E.sync();
    const size_t n = std::min<size_t>(size_t(max_ctas) * 4, E.mac_trace.count);
    SSR_CUDA(cudaMemcpy(out, E.mac_trace.ptr, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  });
}

/* ------------------------------------------------------------- filter sets */

int ssr_filterset_create(ssr_engine* e, int channels, int partitions, ssr_filterset** out)
{
  return guarded([&] {
    require(e && out, "filterset_create: null argument");
    *out = nullptr;
    *out = new ssr_filterset(&e->impl, channels, partitions);
  });
}

void ssr_filterset_destroy(ssr_filterset* fs)
{
  if (!fs) return;
  cudaSetDevice(fs->impl.e->device);
  cudaStreamSynchronize(fs->impl.e->stream);
  delete fs;
}

int ssr_filterset_channels(const ssr_filterset* fs) { return fs ? fs->impl.channels : 0; }
int ssr_filterset_partitions(const ssr_filterset* fs) { return fs ? fs->impl.partitions : 0; }

int ssr_filterset_load_time(ssr_filterset* fs, int first_channel, int channels,
                            const float* data, long frames, long frame_stride, long channel_stride)
{
  return guarded([&] {
    require(fs, "null filterset");
    fs->impl.load_time(first_channel, channels, data, frames, frame_stride, channel_stride);
  });
}

int ssr_filterset_set_partition_time(ssr_filterset* fs, int channel, int partition,
                                     const float* taps, int n)
{
  return guarded([&] {
    require(fs, "null filterset");
    require(n == 0 || taps, "null taps");
    fs->impl.set_partition_time(channel, partition, taps, n);
  });
}

int ssr_filterset_generate(ssr_filterset* fs, int first_channel, int channels, long taps,
                           uint64_t seed, uint64_t channel_id_offset,
                           const float* envelope, float scale)
{
  return guarded([&] {
    require(fs, "null filterset");
    fs->impl.generate(first_channel, channels, taps, seed, channel_id_offset, envelope, scale);
  });
}

void ssr_synth_ir(uint64_t seed, uint64_t channel_id, long taps, const float* envelope,
                  float scale, float* out)
{
  const uint64_t key = ssrk::synth_key(seed, channel_id);
  for (long t = 0; t < taps; ++t)
  {
    const float u = ssrk::synth_uniform(key, uint64_t(t));
    const float env = envelope ? envelope[t] : 1.0f;
    // two separately rounded products, exactly as fwd_sample() on the device
    volatile float ue = u * env;
    out[t] = ue * scale;
  }
}

void ssr_synth_uniform(uint64_t seed, uint64_t stream_id, uint64_t first_index, long n, float* out)
{
  const uint64_t key = ssrk::synth_key(seed, stream_id);
  for (long i = 0; i < n; ++i) out[i] = ssrk::synth_uniform(key, first_index + uint64_t(i));
}

int ssr_filterset_lerp(ssr_filterset* dst, int dch, const ssr_filterset* a, int ach,
                       const ssr_filterset* b, int bch, float t)
{
  return guarded([&] {
    require(dst && a && b, "filterset_lerp: null argument");
    dst->impl.lerp_from(dch, a->impl, ach, b->impl, bch, t);
  });
}

int ssr_filterset_download(ssr_filterset* fs, int channel, int partition, float* sorted_out, int* zero)
{
  return guarded([&] {
    require(fs && sorted_out, "filterset_download: null argument");
    fs->impl.download(channel, partition, sorted_out, zero);
  });
}

int ssr_filterset_upload(ssr_filterset* fs, int channel, int partition, const float* sorted_in, int zero)
{
  return guarded([&] {
    require(fs != nullptr, "filterset_upload: null argument");
    fs->impl.upload(channel, partition, sorted_in, zero);
  });
}

/* ----------------------------------------------------------------- Level 1 */

int ssr_input_create(ssr_engine* e, int partitions, ssr_input** out)
{
  return guarded([&] {
    require(e && out, "input_create: null argument");
    *out = nullptr;
    *out = new ssr_input(&e->impl, partitions);
  });
}

void ssr_input_destroy(ssr_input* in)
{
  if (!in) return;
  cudaSetDevice(in->impl.e->device);
  cudaStreamSynchronize(in->impl.e->stream);
  delete in;
}

int ssr_input_partitions(const ssr_input* in) { return in ? in->impl.P : 0; }

int ssr_input_add_block(ssr_input* in, const float* x)
{
  return guarded([&] { require(in && x, "input_add_block: null argument"); in->impl.add_block(x); });
}

int ssr_input_download(ssr_input* in, int p, float* sorted_out)
{
  return guarded([&] { require(in && sorted_out, "input_download: null argument"); in->impl.download(p, sorted_out); });
}

int ssr_output_create(ssr_input* in, ssr_output_kind kind, ssr_output** out)
{
  return guarded([&] {
    require(in && out, "output_create: null argument");
    *out = nullptr;
    *out = new ssr_output(&in->impl, kind);
  });
}

void ssr_output_destroy(ssr_output* o)
{
  if (!o) return;
  cudaSetDevice(o->impl.e->device);
  cudaStreamSynchronize(o->impl.e->stream);
  delete o;
}

int ssr_output_bind_static(ssr_output* o, const ssr_filterset* fs, int channel)
{
  return guarded([&] {
    require(o && fs, "output_bind_static: null argument");
    require(channel >= 0 && channel < fs->impl.channels, "bad channel");
    o->impl.table.set_static(fs->impl, channel);
  });
}

int ssr_output_set_filter(ssr_output* o, const ssr_filterset* fs, int channel)
{
  return guarded([&] {
    require(o && fs, "output_set_filter: null argument");
    require(o->impl.kind == SSR_OUTPUT_DYNAMIC, "set_filter on a StaticOutput");
    require(channel >= 0 && channel < fs->impl.channels, "bad channel");
    o->impl.table.set_filter(fs->impl, channel);
  });
}

int ssr_output_queues_empty(const ssr_output* o) { return (!o || o->impl.table.queues_empty()) ? 1 : 0; }

int ssr_output_rotate_queues(ssr_output* o)
{
  return guarded([&] { require(o, "null output"); o->impl.table.rotate_queues(); });
}

int ssr_output_convolve(ssr_output* o, float weight, const float** result)
{
  return guarded([&] {
    require(o && result, "output_convolve: null argument");
    *result = o->impl.convolve(weight);
  });
}

/* ----------------------------------------------------------------- Level 2 */

int ssr_renderer_create(ssr_engine* e, const ssr_renderer_config* cfg, ssr_renderer** out)
{
  return guarded([&] {
    require(e && cfg && out, "renderer_create: null argument");
    *out = nullptr;
    *out = new ssr_renderer(&e->impl, *cfg);
  });
}

void ssr_renderer_destroy(ssr_renderer* r) { delete r; }

static void binaural_neutral(Renderer& R, long index)
{
  // neutral filter = Dirac delayed to argmax |h| of channel 0
  // (src/binauralrenderer.h:156-167); may have fewer partitions than the HRTFs
  std::vector<float> impulse(static_cast<size_t>(index) + 1, 0.0f);
  impulse.back() = 1.0f;
  const int P = int(min_partitions(R.e->B, long(impulse.size())));
  R.neutral.reset(new FilterSet(R.e, 1, P));
  R.neutral->load_time(0, 1, impulse.data(), long(impulse.size()), 1, 1);
}

int ssr_renderer_load_hrirs(ssr_renderer* r, const float* data, long frames, int channels, long hrir_size)
{
  return guarded([&] {
    require(r && data, "renderer_load_hrirs: null argument");
    Renderer& R = r->impl;
    require(R.cfg.kind == SSR_RENDERER_BINAURAL, "load_hrirs: not a BinauralRenderer");
    require(R.sources.empty(), "load_hrirs: sources already exist");
    if (channels <= 0 || channels % 2 != 0)
      throw Error(SSR_ERR_FILE, "Number of channels in HRIR file must be a multiple of 2!");
    long size = hrir_size == 0 ? frames : std::min(hrir_size, frames);
    require(size > 0, "load_hrirs: empty HRIR set");
    R.angles = channels / 2;
    R.partitions = int(min_partitions(R.e->B, size));
    R.hrtfs.reset(new FilterSet(R.e, channels, R.partitions));
    R.hrtfs->load_time(0, channels, data, size, channels, 1);
    long index = 0; float best = -1.0f;
    for (long t = 0; t < size; ++t)
    {
      const float v = std::fabs(data[t * channels]);
      if (v > best) { best = v; index = t; }  // std::max_element keeps the first maximum
    }
    binaural_neutral(R, index);
  });
}

int ssr_renderer_load_prefilter(ssr_renderer* r, const float* taps, long frames)
{
  return guarded([&] {
    require(r && taps, "renderer_load_prefilter: null argument");
    Renderer& R = r->impl;
    require(R.cfg.kind == SSR_RENDERER_PREFILTER_BANK, "load_prefilter: not a prefilter bank");
    require(R.sources.empty(), "load_prefilter: inputs already exist");
    require(frames > 0, "load_prefilter: empty pre-filter");
    R.partitions = int(min_partitions(R.e->B, frames));
    R.hrtfs.reset(new FilterSet(R.e, 1, R.partitions));
    R.hrtfs->load_time(0, 1, taps, frames, 1, 1);
  });
}

int ssr_renderer_generate_hrirs(ssr_renderer* r, int channels, long taps, uint64_t seed,
                                const float* envelope, float scale)
{
  return guarded([&] {
    require(r, "null renderer");
    Renderer& R = r->impl;
    require(R.cfg.kind == SSR_RENDERER_BINAURAL, "generate_hrirs: not a BinauralRenderer");
    require(R.sources.empty(), "generate_hrirs: sources already exist");
    if (channels <= 0 || channels % 2 != 0)
      throw Error(SSR_ERR_FILE, "Number of channels in HRIR file must be a multiple of 2!");
    require(taps > 0, "generate_hrirs: empty HRIR set");
    R.angles = channels / 2;
    R.partitions = int(min_partitions(R.e->B, taps));
    R.hrtfs.reset(new FilterSet(R.e, channels, R.partitions));
    R.hrtfs->generate(0, channels, taps, seed, 0, envelope, scale);
    std::vector<float> ch0(static_cast<size_t>(taps), 0.0f);
    ssr_synth_ir(seed, 0, taps, envelope, scale, ch0.data());
    long index = 0; float best = -1.0f;
    for (long t = 0; t < taps; ++t)
    {
      const float v = std::fabs(ch0[t]);
      if (v > best) { best = v; index = t; }
    }
    binaural_neutral(R, index);
  });
}

static std::unique_ptr<FilterSet> source_filterset(Renderer& R, long frames, int channels, int& local_channels, int& first)
{
  if (R.cfg.kind == SSR_RENDERER_GENERIC)
  {
    if (channels != R.M_total)
      throw Error(SSR_ERR_FILE, "apf::load_sndfile(): IR set has " + std::to_string(channels)
          + " channels instead of " + std::to_string(R.M_total) + "!");
    local_channels = R.m_local; first = R.m_first;
  }
  else
  {
    if (channels <= 0 || channels % 2 != 0)
      throw Error(SSR_ERR_FILE, "Number of channels in BRIR file must be a multiple of 2!");
    local_channels = channels; first = 0;
  }
  require(frames > 0, "add_source: empty IR set");
  const int P = int(min_partitions(R.e->B, frames));
  return std::unique_ptr<FilterSet>(new FilterSet(R.e, local_channels, P));
}

int ssr_renderer_add_source(ssr_renderer* r, const float* ir, long frames, int channels, int* id)
{
  return guarded([&] {
    require(r, "null renderer");
    Renderer& R = r->impl;
    int sid;
    if (R.cfg.kind == SSR_RENDERER_BINAURAL || R.cfg.kind == SSR_RENDERER_PREFILTER_BANK)
    {
      require(ir == nullptr, "add_source: this renderer's sources take no IR set");
      sid = R.add_source(nullptr, 0);
    }
    else
    {
      require(ir != nullptr, "add_source: IR set required");
      int local = 0, first = 0;
      auto fs = source_filterset(R, frames, channels, local, first);
      fs->load_time(0, local, ir + first, frames, channels, 1);
      sid = R.add_source(std::move(fs), channels);
    }
    if (id) *id = sid;
  });
}

int ssr_renderer_add_source_synthetic(ssr_renderer* r, long taps, int channels, uint64_t seed,
                                      uint64_t channel_id_offset, const float* envelope,
                                      float scale, int* id)
{
  return guarded([&] {
    require(r, "null renderer");
    Renderer& R = r->impl;
    require(R.cfg.kind != SSR_RENDERER_BINAURAL, "add_source_synthetic: not for BinauralRenderer");
    int local = 0, first = 0;
    auto fs = source_filterset(R, taps, channels, local, first);
    fs->generate(0, local, taps, seed, channel_id_offset + uint64_t(first), envelope, scale);
    const int sid = R.add_source(std::move(fs), channels);
    if (id) *id = sid;
  });
}

int ssr_renderer_rem_source(ssr_renderer* r, int id)
{
  return guarded([&] { require(r, "null renderer"); r->impl.rem_source(id); });
}

int ssr_renderer_sources(const ssr_renderer* r) { return r ? int(r->impl.sources.size()) : 0; }
int ssr_renderer_local_outputs(const ssr_renderer* r) { return r ? r->impl.m_local : 0; }

// Control inputs: callable from any thread, also while another thread is inside
// ssr_renderer_process.  They queue a command (apf::SharedData::operator= ->
// CommandQueue::push, apf/apf/shareddata.h:57-60); the audio thread applies the queue
// at the start of its next block.
static int push_control(ssr_renderer* r, int kind, int id, float a, float b, bool source)
{
  return guarded([&] {
    require(r, "null renderer");
    // (reads the source list: adding / removing sources stays a call of the audio thread's owner)
    if (source) (void)r->impl.find(id);
    if (!r->impl.control.push(ControlQueue::Cmd{kind, id, a, b}))
      throw Error(SSR_ERR_STATE, "ssr_b200: control queue full (4096 commands pending): is the audio thread rendering blocks?");
  });
}

int ssr_source_set_gain(ssr_renderer* r, int id, float gain) { return push_control(r, ControlQueue::SRC_GAIN, id, gain, 0, true); }
int ssr_source_set_mute(ssr_renderer* r, int id, int mute) { return push_control(r, ControlQueue::SRC_MUTE, id, mute ? 1.0f : 0.0f, 0, true); }
int ssr_source_set_position(ssr_renderer* r, int id, float x, float y) { return push_control(r, ControlQueue::SRC_POSITION, id, x, y, true); }
int ssr_source_set_model(ssr_renderer* r, int id, int plane) { return push_control(r, ControlQueue::SRC_MODEL, id, plane ? 1.0f : 0.0f, 0, true); }

int ssr_renderer_set_reference_position(ssr_renderer* r, float x, float y) { return push_control(r, ControlQueue::REF_POSITION, 0, x, y, false); }
int ssr_renderer_set_reference_orientation(ssr_renderer* r, float az) { return push_control(r, ControlQueue::REF_ORIENTATION, 0, az, 0, false); }
int ssr_renderer_set_reference_offset_position(ssr_renderer* r, float x, float y) { return push_control(r, ControlQueue::OFF_POSITION, 0, x, y, false); }
int ssr_renderer_set_reference_offset_orientation(ssr_renderer* r, float az) { return push_control(r, ControlQueue::OFF_ORIENTATION, 0, az, 0, false); }
int ssr_renderer_set_master_volume(ssr_renderer* r, float v) { return push_control(r, ControlQueue::MASTER_VOLUME, 0, v, 0, false); }
int ssr_renderer_set_processing(ssr_renderer* r, int on) { return push_control(r, ControlQueue::PROCESSING, 0, on ? 1.0f : 0.0f, 0, false); }
int ssr_renderer_set_amplitude_reference_distance(ssr_renderer* r, float d) { return push_control(r, ControlQueue::AMP_REF_DISTANCE, 0, d, 0, false); }

int ssr_renderer_get_state(ssr_renderer* r, ssr_renderer_state* out)
{
  return guarded([&] {
    require(r && out, "renderer_get_state: null argument");
    const Renderer& R = r->impl;
    *out = ssr_renderer_state{};
    out->reference_x = R.ref_x; out->reference_y = R.ref_y; out->reference_azimuth = R.ref_az;
    out->reference_offset_x = R.off_x; out->reference_offset_y = R.off_y; out->reference_offset_azimuth = R.off_az;
    out->master_volume = R.master_volume; out->amplitude_reference_distance = R.amplitude_reference_distance;
    out->processing = R.processing ? 1 : 0;
    out->blocks = R.blocks_started;
    out->pending_commands = int(R.control.tail.load() - R.control.head.load());
  });
}

int ssr_source_get_state(ssr_renderer* r, int id, ssr_source_state* out)
{
  return guarded([&] {
    require(r && out, "source_get_state: null argument");
    const Renderer::Source* s = r->impl.find(id);
    *out = ssr_source_state{};
    out->gain = s->gain; out->mute = s->mute ? 1 : 0; out->x = s->px; out->y = s->py; out->plane = s->plane ? 1 : 0;
    out->weighting_factor = s->weighting_factor.get();
  });
}

int ssr_renderer_set_level_metering(ssr_renderer* r, int on)
{
  return guarded([&] { require(r, "null renderer"); r->impl.set_metering(on != 0); });
}

int ssr_renderer_get_levels(ssr_renderer* r, float* source_pre_fader, float* source_levels, float* output_levels)
{
  return guarded([&] {
    require(r, "null renderer");
    r->impl.get_levels(source_pre_fader, source_levels, output_levels);
  });
}

int ssr_renderer_process(ssr_renderer* r, int n, const float* const* in, float* const* out)
{
  return guarded([&] {
    require(r != nullptr, "renderer_process: null argument");
    // a non-root rank of a gather delivers no output blocks: out may be null there
    require(out || r->impl.host_outputs() == 0, "renderer_process: null output array");
    require(in || r->impl.sources.empty(), "renderer_process: null input array");
    r->impl.process(n, in, out);
  });
}

int ssr_renderer_process_batch(ssr_renderer* r, int n, int nblocks, const float* const* in, float* const* out)
{
  return guarded([&] {
    require(r && out, "renderer_process_batch: null argument");
    require(in || r->impl.sources.empty(), "renderer_process_batch: null input array");
    r->impl.process_batch(n, nblocks, in, out);
  });
}

int ssr_renderer_batched_blocks(const ssr_renderer* r, uint64_t* blocks)
{
  return guarded([&] { require(r && blocks, "null argument"); *blocks = r->impl.batched_blocks; });
}

int ssr_renderer_process_device(ssr_renderer* r, const float* d_in, float* d_out)
{
  return guarded([&] {
    require(r && (d_out || r->impl.gather), "renderer_process_device: null argument");
    r->impl.step(d_in, d_out);
  });
}

int ssr_renderer_submit(ssr_renderer* r, int n, const float* const* in)
{
  return guarded([&] { require(r, "null renderer"); r->impl.submit(n, in); });
}

int ssr_renderer_wait(ssr_renderer* r, float* const* out)
{
  return guarded([&] {
    require(r != nullptr, "renderer_wait: null argument");
    require(out || r->impl.host_outputs() == 0, "renderer_wait: null output array");
    r->impl.wait(out);
  });
}

/* ------------------------------------------------------------------ gather */

int ssr_gather_create(ssr_engine* e, int world, int rank, int root, const int* outputs_per_rank,
                      ssr_gather** out)
{
  return guarded([&] {
    require(e && out, "gather_create: null argument");
    *out = nullptr;
    *out = new ssr_gather(&e->impl, world, rank, root, outputs_per_rank);
  });
}

void ssr_gather_destroy(ssr_gather* g) { delete g; }

int ssr_gather_export(ssr_gather* g, void* handle)
{
  return guarded([&] { require(g && handle, "gather_export: null argument"); g->impl.export_handle(handle); });
}

int ssr_gather_import(ssr_gather* g, const void* handle)
{
  return guarded([&] { require(g && handle, "gather_import: null argument"); g->impl.import_handle(handle); });
}

int ssr_gather_connect_local(ssr_gather* g, ssr_gather* root)
{
  return guarded([&] {
    require(g && root, "gather_connect_local: null argument");
    require(root->impl.is_root() && root->impl.base, "gather_connect_local: second argument is not a root gather");
    require(root->impl.M_total == g->impl.M_total && root->impl.world == g->impl.world
        && root->impl.e->B == g->impl.e->B, "gather_connect_local: layouts differ");
    g->impl.import_pointer(root->impl.base, root->impl.e->device);
  });
}

int ssr_renderer_bind_gather(ssr_renderer* r, ssr_gather* g)
{
  return guarded([&] { require(r, "null renderer"); r->impl.bind_gather(g ? &g->impl : nullptr); });
}

int ssr_gather_wait(ssr_gather* g, const float** d_block)
{
  return guarded([&] {
    require(g && d_block, "gather_wait: null argument");
    *d_block = g->impl.wait();
  });
}

int ssr_gather_release(ssr_gather* g)
{
  return guarded([&] { require(g, "null gather"); g->impl.release(); });
}

int ssr_gather_status(ssr_gather* g)
{
  int flag = -1;
  int rc = guarded([&] { require(g, "null gather"); flag = g->impl.status(); });
  if (rc != SSR_OK) return rc;
  if (flag) { g_error = "ssr_b200: gather: a wait on a peer rank timed out"; return SSR_ERR_STATE; }
  return SSR_OK;
}

int ssr_gather_total_outputs(const ssr_gather* g) { return g ? g->impl.M_total : 0; }

/* ------------------------------------------------------------- hostgather */

int ssr_hostgather_create(ssr_engine* e, int world, int rank, int root, const int* outputs_per_rank,
                          const char* shm_name, ssr_hostgather** out)
{
  return guarded([&] {
    require(e && out && shm_name, "hostgather_create: null argument");
    *out = nullptr;
    *out = new ssr_hostgather(&e->impl, world, rank, root, outputs_per_rank, shm_name);
  });
}

void ssr_hostgather_destroy(ssr_hostgather* g) { delete g; }

int ssr_renderer_bind_hostgather(ssr_renderer* r, ssr_hostgather* g)
{
  return guarded([&] { require(r, "null renderer"); r->impl.bind_hostgather(g ? &g->impl : nullptr); });
}

int ssr_hostgather_total_outputs(const ssr_hostgather* g) { return g ? g->impl.M_total : 0; }

int ssr_hostgather_slot(ssr_hostgather* g, int slot, float** host_blocks)
{
  return guarded([&] {
    require(g && host_blocks, "hostgather_slot: null argument");
    require(slot == 0 || slot == 1, "hostgather_slot: slot must be 0 or 1");
    *host_blocks = g->impl.slot_ptr(slot);
  });
}

int ssr_hostgather_next_slot(const ssr_hostgather* g) { return g ? int(g->impl.epoch & 1) : 0; }

}  // extern "C"
