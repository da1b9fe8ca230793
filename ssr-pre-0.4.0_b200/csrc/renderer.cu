// renderer.cu -- Level 2: the per-block step of the Generic / Binaural / BRS
// renderers, batched into  forward FFT -> MAC -> inverse FFT + crossfade + sum.
//
// Host logic (weights, crossfade modes, filter index, staggered filter switch)
// follows the reference exactly; arithmetic is moved to the kernels:
//   src/rendererbase.h:386-407      weighting_factor
//   src/genericrenderer.h:122-180   Generic source phase + RenderFunction::select
//   src/brsrenderer.h:141-210       BRS source phase
//   src/binauralrenderer.h:261-389  Binaural source phase
//   apf/apf/combine_channels.h:89-130, 292-335, 366-403  crossfade combine
// The reference evaluates every (source, output) pair in the time domain and adds
// the results; here sources are summed in the frequency domain per accumulator
// kind (constant / fade-out / fade-in) so only <= 3 inverse FFTs per output run.
#include "engine.hpp"

#include <algorithm>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace ssr_b200
{

using namespace ssrk;

static constexpr int kGenericMT = 4;

Renderer::Renderer(Engine* e_, const ssr_renderer_config& c)
  : e(e_), cfg(c)
{
  e->use_device();
  if (cfg.kind == SSR_RENDERER_GENERIC)
  {
    if (cfg.outputs <= 0) throw Error(SSR_ERR_INVALID, "ssr_b200: GenericRenderer needs outputs > 0");
    M_total = cfg.outputs;
    m_first = cfg.output_first;
    m_local = cfg.output_count > 0 ? cfg.output_count : M_total - m_first;
    if (m_first < 0 || m_local <= 0 || m_first + m_local > M_total)
      throw Error(SSR_ERR_INVALID, "ssr_b200: bad output shard");
  }
  else if (cfg.kind == SSR_RENDERER_PREFILTER_BANK)
  {
    M_total = 0; m_first = 0; m_local = 0;  // one output per input, grows with add_source
  }
  else
  {
    if (cfg.outputs != 0 && cfg.outputs != 2)
      throw Error(SSR_ERR_INVALID, "ssr_b200: Binaural/BRS renderers have exactly 2 outputs");
    M_total = 2; m_first = 0; m_local = 2;
  }
  // apf::math::dB2linear (src/rendererbase.h:234-235)
  mvc = float(std::pow(10.0, double(cfg.master_volume_correction_db) / 20.0));
  if (std::getenv("SSR_B200_NO_GRAPHS")) use_graphs = false;  // experiments
  if (const char* v = std::getenv("SSR_B200_CTL_UPLOAD")) ctl_by_kernel = std::string(v) != "copy";   // A/B
  if (const char* v = std::getenv("SSR_B200_DYN_BLOCK_OVERLAP")) dyn_block_overlap = std::atoi(v) != 0;   // A/B
  if (const char* v = std::getenv("SSR_B200_ZEROCOPY_OUT")) zero_copy = std::atoi(v) != 0;
  if (const char* v = std::getenv("SSR_B200_ZEROCOPY_IN")) zero_copy_in = std::atoi(v) != 0;
  static_mac.MT = kGenericMT;
  for (auto& d : d_out) d.resize(size_t(std::max(m_local, 1)) * e->B);
  h_out.resize(size_t(2) * std::max(m_local, 1) * e->B);
}

Renderer::~Renderer()
{
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  for (auto ev : done) if (ev) cudaEventDestroy(ev);
  if (calib.ev) cudaEventDestroy(calib.ev);
  for (auto ev : dyn_copied) if (ev) cudaEventDestroy(ev);
  if (s_in) { cudaStreamSynchronize(s_in); cudaStreamSynchronize(s_out); }
  for (auto ev : ev_in) if (ev) cudaEventDestroy(ev);
  for (auto ev : ev_k) if (ev) cudaEventDestroy(ev);
  if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_out); }
  drop_graphs();
  prof.report("s0 source logic, s1 plan, s2 inv+arena, s3 fwd table, s4 launches");
}

Renderer::Source* Renderer::find_or_null(int id)
{
  for (auto& s : sources) if (s->id == id) return s.get();
  return nullptr;
}

Renderer::Source* Renderer::find(int id)
{
  if (Source* s = find_or_null(id)) return s;
  throw Error(SSR_ERR_INVALID, "ssr_b200: no such source");
}

bool ControlQueue::push(const Cmd& c)
{
  std::lock_guard<std::mutex> lock(producers);
  for (int tries = 0; ; ++tries)
  {
    const uint32_t t = tail.load(std::memory_order_relaxed);
    if (t - head.load(std::memory_order_acquire) < kCapacity)
    {
      ring[t % kCapacity] = c;
      tail.store(t + 1, std::memory_order_release);
      return true;
    }
    // "if the FIFO is full: retry, retry, ..." with usleep(50) (apf/apf/commandqueue.h:203-209)
    if (tries > 4000) return false;
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}

// CommandQueue::process_commands() (apf/apf/commandqueue.h:141-153) at the start of the
// audio block (apf/apf/mimoprocessor.h:374): the SetCommands execute in order
void Renderer::apply_controls()
{
  ++blocks_started;
  control.drain([this](const ControlQueue::Cmd& c)
  {
    Source* s = nullptr;
    if (c.kind <= ControlQueue::SRC_MODEL && !(s = find_or_null(c.id))) return;  // removed meanwhile
    switch (c.kind)
    {
      case ControlQueue::SRC_GAIN: s->gain = c.a; break;
      case ControlQueue::SRC_MUTE: s->mute = c.a != 0.0f; break;
      case ControlQueue::SRC_POSITION: s->px = c.a; s->py = c.b; break;
      case ControlQueue::SRC_MODEL: s->plane = c.a != 0.0f; break;
      case ControlQueue::REF_POSITION: ref_x = c.a; ref_y = c.b; break;
      case ControlQueue::REF_ORIENTATION: ref_az = c.a; break;
      case ControlQueue::OFF_POSITION: off_x = c.a; off_y = c.b; break;
      case ControlQueue::OFF_ORIENTATION: off_az = c.a; break;
      case ControlQueue::MASTER_VOLUME: master_volume = c.a; break;
      case ControlQueue::PROCESSING: processing = c.a != 0.0f; break;
      case ControlQueue::AMP_REF_DISTANCE: amplitude_reference_distance = c.a; break;
    }
  });
}

int Renderer::add_source(std::unique_ptr<FilterSet> irs, int channels)
{
  e->use_device();
  e->sync();
  auto s = std::unique_ptr<Source>(new Source);
  s->id = next_id++;
  if (cfg.kind == SSR_RENDERER_GENERIC)
  {
    s->irs = std::move(irs);
    // Input(block, min_partitions(block, frames)) (src/genericrenderer.h:109-110)
    s->input.reset(new Input(e, s->irs->partitions));
    (void)channels;
  }
  else if (cfg.kind == SSR_RENDERER_BRS)
  {
    s->brtf = std::move(irs);
    s->angles = channels / 2;
    s->input.reset(new Input(e, s->brtf->partitions));
    for (int i = 0; i < 2; ++i) s->hist[i].assign(size_t(s->brtf->partitions) + 1, DynRingEntry{nullptr, nullptr, 0, 0});
  }
  else if (cfg.kind == SSR_RENDERER_PREFILTER_BANK)
  {
    if (!hrtfs) throw Error(SSR_ERR_STATE, "ssr_b200: prefilter bank: load the pre-filter before adding inputs");
    // StaticConvolver(*_pre_filter) per input (src/wfsrenderer.h:111-113)
    s->input.reset(new Input(e, partitions));
    M_total = m_local = int(sources.size()) + 1;
    for (auto& d : d_out) d.resize(size_t(m_local) * e->B);
    h_out.resize(size_t(2) * m_local * e->B);
    invalidate_inv();
  }
  else
  {
    if (!hrtfs) throw Error(SSR_ERR_STATE, "ssr_b200: BinauralRenderer: load HRIRs before adding sources");
    s->input.reset(new Input(e, partitions));
    for (int i = 0; i < 2; ++i) s->hist[i].assign(size_t(partitions) + 1, DynRingEntry{nullptr, nullptr, 0, 0});
    s->temp.resize(2);
    s->bw = BlockParameter<float>(0.0f);  // _weight(0.0f) (src/binauralrenderer.h:245)
  }
  sources.push_back(std::move(s));
  if (metering) { d_levels.release(); (void)level_ptr(0); }
  invalidate_inv();
  static_dirty = true;
  fwd_cache.clear();
  return sources.back()->id;
}

void Renderer::rem_source(int id)
{
  e->use_device();
  e->sync();
  for (size_t i = 0; i < sources.size(); ++i)
    if (sources[i]->id == id)
    {
      sources.erase(sources.begin() + i);
      if (cfg.kind == SSR_RENDERER_PREFILTER_BANK) { M_total = m_local = int(sources.size()); invalidate_inv(); }
      static_dirty = true;
      fwd_cache.clear();
      return;
    }
  throw Error(SSR_ERR_INVALID, "ssr_b200: no such source");
}

// level meters (apf::enable_queries, src/rendererbase.h:431-437, 468-474)
float* Renderer::level_ptr(int index)
{
  if (!metering) return nullptr;
  const size_t need = sources.size() + size_t(std::max(m_local, 1)) + 8;
  if (d_levels.count < need)
  {
    e->sync();
    d_levels.resize(need + 64);
    SSR_CUDA(cudaMemsetAsync(d_levels.ptr, 0, d_levels.count * sizeof(float), e->stream));
  }
  return d_levels.ptr + index;
}

void Renderer::set_metering(bool on)
{
  if (on == metering) return;
  e->sync();
  metering = on;
  if (on) (void)level_ptr(0);
  fwd_cache.clear();
  invalidate_inv();
}

void Renderer::get_levels(float* pre_fader, float* levels, float* outputs)
{
  if (!metering) throw Error(SSR_ERR_STATE, "ssr_b200: level metering is off");
  e->use_device();
  e->sync();
  const int N = int(sources.size());
  std::vector<float> tmp(size_t(N) + m_local);
  if (!tmp.empty())
    SSR_CUDA(cudaMemcpy(tmp.data(), level_ptr(0), tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (int n = 0; n < N; ++n)
  {
    if (pre_fader) pre_fader[n] = tmp[n];
    // _level = _pre_fader_level * weighting_factor (src/rendererbase.h:434)
    if (levels) levels[n] = tmp[n] * sources[n]->weighting_factor.get();
  }
  for (int m = 0; m < m_local; ++m) if (outputs) outputs[m] = tmp[N + m];
}

// RendererBase::Source::Process (src/rendererbase.h:386-407)
void Renderer::base_weight(Source& s)
{
  if (!processing || s.mute) s.weighting_factor.assign(0.0f);
  else
  {
    float w = s.gain;
    w *= master_volume;
    w *= mvc;
    s.weighting_factor.assign(w);
  }
}

void Renderer::build_fwd(const float* d_input)
{
  const int N = int(sources.size());
  for (auto& c : fwd_cache)
    if (c.first == d_input) { fwd = c.second.get(); return; }
  if (fwd_cache.size() >= 32) { e->sync(); fwd_cache.clear(); }
  fwd_cache.emplace_back(d_input, std::unique_ptr<Table<FwdJob>>(new Table<FwdJob>()));
  fwd = fwd_cache.back().second.get();
  fwd->host.resize(N);
  for (int n = 0; n < N; ++n)
  {
    fwd->host[n] = sources[n]->input->make_job(d_input + size_t(n) * e->B);
    fwd->host[n].level = level_ptr(n);
  }
  fwd->upload(e->stream);
}

static Mode mode_from_weight(const BlockParameter<float>& f)
{
  // GenericRenderer::RenderFunction::select (src/genericrenderer.h:161-180)
  if (f.both_eq(0.0f)) return NOTHING;
  if (f.old() == 0.0f) return FADE_IN;
  if (f.get() == 0.0f) return FADE_OUT;
  if (!f.changed()) return CONSTANT;
  return CHANGE;
}

// Share calibration of the bulk-copy MAC (see Renderer::Calibration).  Called once per block
// of the Generic renderer, before the block's kernels are queued.
void Renderer::calibrate(int kind_mask)
{
  constexpr int kSettle = 8, kMeasure = 16, kRounds = 2;
  if (!e->mac_calibrate || !e->mac_tma() || static_mac.MT != kGenericMT || static_mac.ncta < 2
      || !e->mac_trace.ptr || static_mac.ncta > int(e->mac_trace.count / 4) || calib.state == 3)
    return;
  const bool steady = kind_mask == 1;   // one MAC launch per block, constant weights
  if (calib.state == 2)
  {
    if (cudaEventQuery(calib.ev) != cudaSuccess) { cudaGetLastError(); return; }
    // rates: terms of the share / busy time accumulated over the measured launches
    const int C = static_mac.ncta;
    std::vector<double> rate(size_t(C), 0.0);
    bool ok = true;
    for (int c = 0; c < C && ok; ++c)
    {
      long terms = 0;
      for (int i = static_mac.cta_first.host[size_t(c)]; i < static_mac.cta_first.host[size_t(c) + 1]; ++i)
        terms += static_mac.work.host[size_t(i)].y - static_mac.work.host[size_t(i)].x;
      const double busy = double(calib.host.ptr[size_t(c) * 4 + 3]);
      ok = terms > 0 && busy > 0.0;
      if (ok) rate[size_t(c)] = double(terms) / busy;
    }
    double mean = 0.0;
    for (double r : rate) mean += r / double(C);
    for (int c = 0; c < C && ok; ++c) ok = rate[size_t(c)] > 0.5 * mean && rate[size_t(c)] < 2.0 * mean;  // sane
    if (ok)
    {
      e->sync();   // kernels in flight still read the old work tables
      calib.weights = rate;
      static_mac.make_work(e->mac_slots(static_mac.MT), e->split_cost(static_mac.MT), e->mac_waves, &calib.weights);
      static_mac.upload(e->stream);
      e->ensure_partials(size_t(3) * static_mac.npartials * static_mac.MT);
      invalidate_inv();
      inv_batch_key = -1;
      ++calib.rounds;
    }
    calib.state = (calib.rounds >= kRounds || !ok) ? 3 : 0;
    calib.steady = 0;
    return;
  }
  if (!steady) { if (calib.state == 1) calib.state = 0; calib.steady = 0; return; }
  if (calib.state == 0)
  {
    if (++calib.steady < kSettle) return;
    // start measuring with this block's launch: clear the accumulated busy times (stream-ordered)
    SSR_CUDA(cudaMemsetAsync(e->mac_trace.ptr, 0, size_t(static_mac.ncta) * 4 * sizeof(unsigned long long), e->stream));
    calib.state = 1;
    calib.measured = 0;
    return;
  }
  if (calib.state == 1 && ++calib.measured > kMeasure)
  {
    // the launches of the last kMeasure blocks are queued before this copy
    if (calib.host.count < size_t(static_mac.ncta) * 4) calib.host.resize(size_t(static_mac.ncta) * 4);
    if (!calib.ev) SSR_CUDA(cudaEventCreateWithFlags(&calib.ev, cudaEventDisableTiming));
    SSR_CUDA(cudaMemcpyAsync(calib.host.ptr, e->mac_trace.ptr, size_t(static_mac.ncta) * 4 * sizeof(unsigned long long),
          cudaMemcpyDeviceToHost, e->stream));
    SSR_CUDA(cudaEventRecord(calib.ev, e->stream));
    calib.state = 2;
  }
}

Renderer::StepDesc Renderer::prepare_generic(const float* d_input, float* d_output)
{
  InvPlan& inv = inv_plans[cur_slot];   // one set of inverse tables per output buffer
  int& inv_key = inv_keys[cur_slot];
  StepDesc d;
  d.d_input = d_input; d.d_output = d_output;
  const int N = int(sources.size());
  const int B = e->B;
  wtab_host.assign(size_t(3) * std::max(N, 1), 0.0f);
  bool kind_active[3] = {false, false, false};
  uint64_t active_parts[3] = {0, 0, 0};
  for (int n = 0; n < N; ++n)
  {
    Source& s = *sources[n];
    if (tail_level && !s.primed)
    {
      // (non-uniform partitioning, tail level: what this renderer adds sounds 2 K blocks after the
      // source started, when the reference's initial fade-in is long over)
      base_weight(s);
      s.w.assign(s.weighting_factor.get());
      s.primed = true;
    }
    base_weight(s);
    s.w.assign(s.weighting_factor.get());  // src/genericrenderer.h:124
    s.mode = mode_from_weight(s.w);
    const uint64_t parts = uint64_t(s.irs->valid.empty() ? 0 : s.irs->valid[0]);
    if (s.mode == CONSTANT) { wtab_host[n] = s.w.get(); kind_active[0] = true; active_parts[0] += parts; }
    if (s.mode == CHANGE || s.mode == FADE_OUT) { wtab_host[N + n] = s.w.old(); kind_active[1] = true; active_parts[1] += parts; }
    if (s.mode == CHANGE || s.mode == FADE_IN) { wtab_host[2 * N + n] = s.w.get(); kind_active[2] = true; active_parts[2] += parts; }
  }

  if (static_dirty)
  {
    e->sync();
    static_mac.clear();
    const float4* zero = reinterpret_cast<const float4*>(e->zero_part.ptr);
    // order of a group's (source, partition) terms in the table (a CTA's share is a contiguous
    // run of it): 0 = source-major, 1 = partition-major, 2 = pseudo-random (experiments)
    int term_order = 0;
    if (const char* v = std::getenv("SSR_B200_TERM_ORDER")) term_order = std::atoi(v);
    std::vector<std::pair<int, int>> np_list;
    int Pmax = 0;
    for (int n = 0; n < N; ++n) Pmax = std::max(Pmax, sources[n]->irs->partitions);
    if (term_order == 1)
    {
      for (int p = 0; p < Pmax; ++p)
        for (int n = 0; n < N; ++n) if (p < sources[n]->irs->partitions) np_list.emplace_back(n, p);
    }
    else
    {
      for (int n = 0; n < N; ++n)
        for (int p = 0; p < sources[n]->irs->partitions; ++p) np_list.emplace_back(n, p);
      if (term_order == 2)
      {
        uint64_t st = 0x9E3779B97F4A7C15ull;
        for (size_t i = np_list.size(); i > 1; --i)
        {
          st = st * 6364136223846793005ull + 1442695040888963407ull;
          std::swap(np_list[i - 1], np_list[size_t((st >> 33) % i)]);
        }
      }
    }
    for (int m0 = 0; m0 < m_local; m0 += kGenericMT)
    {
      static_mac.begin_group();
      for (auto& np : np_list)
      {
        const int n = np.first, p = np.second;
        Source& s = *sources[n];
        const float4* h[kGenericMT];
        bool any = false;
        for (int j = 0; j < kGenericMT; ++j)
        {
          const float2* ptr = (m0 + j < m_local) ? s.irs->part_or_null(m0 + j, p) : nullptr;
          h[j] = ptr ? reinterpret_cast<const float4*>(ptr) : zero;
          any |= (ptr != nullptr);
        }
        if (any) static_mac.add_term(s.input->id, p, n, h);
      }
      static_mac.end_group();
    }
    calib.reset();
    static_mac.make_work(e->mac_slots(static_mac.MT), e->split_cost(static_mac.MT), e->mac_waves);
    static_mac.upload(e->stream);
    e->ensure_partials(size_t(3) * static_mac.npartials * static_mac.MT);
    static_dirty = false;
    invalidate_inv();
  }

  const int key = (kind_active[0] ? 1 : 0) | (kind_active[1] ? 2 : 0) | (kind_active[2] ? 4 : 0);
  calibrate(key);   // may re-cut the MAC shares (then the inverse tables below are rebuilt)
  if (key != inv_key || (inv.jobs.host.size() && inv.jobs.host[0].out != d_output))
  {
    inv.clear();
    const float scale = 1.0f / float(2 * B);
    for (int m = 0; m < m_local; ++m)
    {
      const int g = m / kGenericMT, row = m % kGenericMT;
      const auto& grp = static_mac.groups[g];
      InvJob job{int(inv.entries.host.size()), 0, d_output + size_t(m) * B, level_ptr(int(sources.size()) + m)};
      for (int k = 0; k < 3; ++k)
      {
        if (!kind_active[k]) continue;
        inv.entries.host.push_back({k * static_mac.npartials + grp.partial_first, grp.nsplit,
            kGenericMT, row, k, scale});
        job.n_entries++;
      }
      inv.jobs.host.push_back(job);
    }
    inv.upload(e->stream);
    inv_key = key;
  }

  build_fwd(d_input);
  e->upload_weights(wtab_host.data(), wtab_host.size());
  d.N = N;
  d.kind_mask = key;
  for (int k = 0; k < 3; ++k)
  {
    // H partitions, X partitions; + m_local accumulators per pass
    d.hparts[k] = active_parts[k] * uint64_t(m_local);
    d.xparts[k] = active_parts[k];
  }
  d.nrows = m_local;
  d.graphable = gather == nullptr && hostgather == nullptr;  // a gather's kernel arguments change every block
  return d;
}

// apf::math::wrap<float> (apf/apf/math.h:136-160)
static float fwrap(float x, float full)
{
  x = std::fmod(x, full);
  if (x < 0) x += full;
  return x;
}

// Binaural / BRS.  The host keeps the reference's per-source decisions (weights,
// filter index, crossfade mode, queues_empty) and uploads ~80 bytes per source;
// the (source, partition) term list of the staggered filter switch is generated
// inside the MAC kernel from a per-(source, ear) ring of "current filter by block"
// (ssrk::DynTables).  Every launch has a fixed shape, so the block replays as a
// CUDA graph like the static renderers.
Renderer::StepDesc Renderer::prepare_dynamic(const float* d_input, float* d_output)
{
  InvPlan& inv = inv_plans[cur_slot];   // one set of inverse tables per output buffer
  int& inv_key = inv_keys[cur_slot];
  StepDesc d;
  d.d_input = d_input; d.d_output = d_output;
  const int N = int(sources.size());
  const int B = e->B;
  const bool binaural = cfg.kind == SSR_RENDERER_BINAURAL;
  prof.start();

  if (static_dirty)
  {
    // source list changed: re-lay the device ring from the hosts' histories
    e->sync();
    dyn_Pmax = 1;
    for (auto& s : sources) dyn_Pmax = std::max(dyn_Pmax, s->input->P);
    dyn_R = dyn_Pmax + 1;
    std::vector<DynRingEntry> ring(size_t(std::max(N, 1)) * 2 * dyn_R, DynRingEntry{nullptr, nullptr, 0, 0});
    for (int n = 0; n < N; ++n)
    {
      Source& s = *sources[n];
      const long long Hn = s.input->P + 1;
      for (int ear = 0; ear < 2; ++ear)
        for (long long tau = dyn_b - (Hn - 1); tau <= dyn_b; ++tau)
        {
          if (tau < 0) continue;
          ring[(size_t(n) * 2 + ear) * dyn_R + size_t(tau % dyn_R)] = s.hist[ear][size_t(tau % Hn)];
        }
    }
    dyn_ring.resize(ring.size());
    SSR_CUDA(cudaMemcpy(dyn_ring.ptr, ring.data(), ring.size() * sizeof(DynRingEntry), cudaMemcpyHostToDevice));
    // CTAs of the generated-terms MAC: what is resident at once -- next to the forward
    // kernel's CTAs when that runs under the MAC (one CTA per source, each takes the
    // registers of one MAC CTA on its SM)
    dyn_G = e->dyn_mac_slots();
    if (e->dyn_tma())
      // one CTA per SM, each with an equal run of (source, partition) slots and ONE set of
      // [kind][ear] partial spectra (mac_dyn_tma_kernel)
      dyn_G = int(std::max<long long>(1, std::min<long long>(e->sm_count, (long long)N * dyn_Pmax)));
    else if (e->fwd_overlap) dyn_G = std::max(e->sm_count, dyn_G - std::min(N, e->sm_count));
    if (const char* v = std::getenv("SSR_B200_DYN_CTAS")) dyn_G = std::max(1, std::atoi(v));  // experiments
    e->ensure_partials(size_t(dyn_G) * (e->dyn_tma() ? 6 : 2));
    static_dirty = false;
    invalidate_inv();
  }

  // layout of the per-block control upload
  const size_t off_lists = sizeof(DynCtl);
  const size_t off_input = off_lists + size_t(3) * N * sizeof(int);
  const size_t off_parts = off_input + size_t(N) * sizeof(int);
  const size_t off_w = off_parts + size_t(N) * sizeof(int);
  size_t off_upd = off_w + size_t(3) * N * sizeof(float);
  off_upd = (off_upd + 15) & ~size_t(15);
  const size_t total = off_upd + size_t(2) * N * sizeof(DynRingEntry);
  dyn_slot ^= 1;
  if (ctl_by_kernel)
  {
    if (!dyn_ack.ptr)
    {
      dyn_ack.resize(2); dyn_ack.ptr[0] = dyn_ack.ptr[1] = 0;
      dyn_upload_count.resize(2);
      SSR_CUDA(cudaMemsetAsync(dyn_upload_count.ptr, 0, 2 * sizeof(unsigned long long), e->stream));
    }
    // the upload kernel that last read this staging slot has run (normally long ago)
    const volatile unsigned long long* ack = dyn_ack.ptr + dyn_slot;
    if (*ack < dyn_uploads_issued[dyn_slot])
    {
      const auto t0 = std::chrono::steady_clock::now();
      while (*ack < dyn_uploads_issued[dyn_slot])
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20))
          throw Error(SSR_ERR_STATE, "ssr_b200: control upload of an earlier block never ran");
    }
  }
  else
  {
    if (dyn_copied[dyn_slot]) SSR_CUDA(cudaEventSynchronize(dyn_copied[dyn_slot]));
    else SSR_CUDA(cudaEventCreateWithFlags(&dyn_copied[dyn_slot], cudaEventDisableTiming));
  }
  if (dyn_stage[dyn_slot].count < total + 16)
  {
    // (a slot that grows may still be read by an upload in flight)
    if (ctl_by_kernel) e->sync();
    dyn_stage[dyn_slot].resize(total + total / 2 + 256);
  }
  DeviceArray<unsigned char>& dyn_dev_cur = dyn_dev[dyn_slot];
  if (dyn_dev_cur.count < total + 16) { e->sync(); dyn_dev_cur.resize(total + total / 2 + 256); }
  unsigned char* st = dyn_stage[dyn_slot].ptr;
  DynCtl* ctl = reinterpret_cast<DynCtl*>(st);
  int* lists = reinterpret_cast<int*>(st + off_lists);
  int* src_input = reinterpret_cast<int*>(st + off_input);
  int* src_parts = reinterpret_cast<int*>(st + off_parts);
  float* wtab = reinterpret_cast<float*>(st + off_w);
  DynRingEntry* upd = reinterpret_cast<DynRingEntry*>(st + off_upd);
  std::memset(wtab, 0, size_t(3) * N * sizeof(float));

  lerp_tab.host.clear();
  const long long b = ++dyn_b;   // block index of the state after this block's switch
  int nk[3] = {0, 0, 0};
  auto select = [&](int n, int kind, float w)
  {
    lists[kind * N + nk[kind]++] = n;
    wtab[kind * N + n] = w;
  };

  for (int n = 0; n < N; ++n)
  {
    Source& s = *sources[n];
    const int P = s.input->P;
    base_weight(s);
    bool filter_changed;
    if (!binaural)
    {
      // src/brsrenderer.h:141-184
      s.bw.assign(s.weighting_factor.get());
      const float azi = ref_az;
      s.index.assign(long(size_t(fwrap((azi - 90.0f) * float(s.angles) / 360.0f + 0.5f,
                float(s.angles)))));
      filter_changed = s.index.changed();
    }
    else
    {
      // src/binauralrenderer.h:261-314
      float interp_factor = 0.0f, weight = 0.0f;
      const float rx = ref_x + off_x, ry = ref_y + off_y;
      const float ref_ori = ref_az + off_az;
      const float dx = s.px - rx, dy = s.py - ry;
      if (s.weighting_factor.get() != 0)
      {
        weight = 1;
        if (s.plane) weight *= 0.5f / amplitude_reference_distance;
        else
        {
          float dist = std::sqrt(dx * dx + dy * dy);
          if (dist < 0.5f) interp_factor = 1.0f - 2 * dist;
          dist = std::max(dist, 0.5f);
          weight *= 0.5f / dist;
        }
        weight *= s.weighting_factor.get();
      }
      s.interp.assign(interp_factor);
      s.bw.assign(weight);
      const float fangles = float(angles);
      // Position::orientation (src/position.cpp:76-79)
      const float src_az = std::atan2(dy, dx) / (float(M_PI) / 180.0f);
      const float rel = src_az - ref_ori;
      s.index.assign(long(size_t(fwrap(rel * fangles / 360.0f + 0.5f, fangles))));
      filter_changed = s.index.changed() || s.interp.changed();
    }
    const BlockParameter<float>& w = s.bw;
    // Output::queues_empty() (convolver.h:666-677) before this block's rotate: the
    // last partition of a filter installed in block b0 becomes active in block
    // b0 + P - 1, so the queues are empty from block b0 + P on
    const bool queues_empty = P < 2 || !s.has_switched || b >= s.last_switch + P;
    Mode mode;
    if (w.both_eq(0.0f)) mode = NOTHING;
    else if (queues_empty && !w.changed() && !filter_changed) mode = CONSTANT;
    else if (binaural)
    {
      // fade_out is tested before fade_in (src/binauralrenderer.h:332-339)
      if (w.get() == 0.0f) mode = FADE_OUT;
      else if (w.old() == 0.0f) mode = FADE_IN;
      else mode = CHANGE;
    }
    else
    {
      // fade_in before fade_out (src/brsrenderer.h:171-178)
      if (w.old() == 0.0f) mode = FADE_IN;
      else if (w.get() == 0.0f) mode = FADE_OUT;
      else mode = CHANGE;
    }
    s.mode = mode;

    // old state (convolve_and_more(old weight)) = the tables of block b - 1;
    // rotate_queues() is implicit in the block index; then the switch
    if (mode != NOTHING && mode != FADE_IN && mode != CONSTANT) select(n, 1, w.old());
    if (filter_changed)
    {
      for (int i = 0; i < 2; ++i)
      {
        if (!binaural)
        {
          const int hch = int(2 * s.index.get() + i);
          s.latest[i] = DynRingEntry{reinterpret_cast<const float4*>(s.brtf->part(hch, 0)),
              s.brtf->device_flags_row(hch), s.brtf->partitions, 0};
        }
        else
        {
          const int hch = int(2 * s.index.get() + i);
          if (s.interp.get() == 0.0f)
            s.latest[i] = DynRingEntry{reinterpret_cast<const float4*>(hrtfs->part(hch, 0)),
                hrtfs->device_flags_row(hch), hrtfs->partitions, 0};
          else
          {
            // Dirac interpolation into a temporary filter
            // (src/binauralrenderer.h:365-379).  A temporary that is still
            // referenced by an active or queued partition is not overwritten
            // (the reference aliases them); a free one from
            // the pool is used instead.
            FilterSet* tmp = nullptr;
            for (auto& t : s.temp[i])
            {
              const float4* lo = reinterpret_cast<const float4*>(t->data.ptr);
              const float4* hi = reinterpret_cast<const float4*>(t->data.ptr + t->data.count);
              bool used = false;
              for (auto& h : s.hist[i]) used |= (h.base >= lo && h.base < hi);
              if (!used) { tmp = t.get(); break; }
            }
            if (!tmp)
            {
              s.temp[i].emplace_back(new FilterSet(e, 1, partitions));
              tmp = s.temp[i].back().get();
            }
            tmp->lerp_plan(0, *hrtfs, hch, *neutral, 0, s.interp.get(), lerp_tab.host);
            s.latest[i] = DynRingEntry{reinterpret_cast<const float4*>(tmp->part(0, 0)),
                tmp->device_flags_row(0), tmp->partitions, 0};
          }
        }
      }
      s.has_switched = true;
      s.last_switch = b;
    }
    for (int i = 0; i < 2; ++i)
    {
      s.hist[i][size_t(b % (P + 1))] = s.latest[i];
      upd[n * 2 + i] = s.latest[i];
    }
    src_input[n] = s.input->id;
    src_parts[n] = P;
    if (mode == CONSTANT) select(n, 0, w.get());
    else if (mode == CHANGE || mode == FADE_IN) select(n, 2, w.get());
  }
  if (!lerp_tab.host.empty())
  {
    // this block's Dirac interpolations (src/binauralrenderer.h:365-379) in ONE stream-ordered
    // launch: preallocated pinned + device job table, no cudaMalloc / sync on the audio path
    lerp_tab.upload(e->stream);
    lerp_kernel<<<int(lerp_tab.host.size()), kThreads, 0, e->stream>>>(lerp_tab.device(), B / 2);
    SSR_CUDA(cudaGetLastError());
    e->count_launch();
  }
  ctl->b = b;
  ctl->n[0] = nk[0]; ctl->n[1] = nk[1]; ctl->n[2] = nk[2]; ctl->pad = 0;
  prof.lap(0);

  unsigned char* dev_ctl = dyn_dev_cur.ptr;
  if (ctl_by_kernel)
  {
    void* mapped = nullptr;   // the device's view of the staging slot (the same address under UVA)
    SSR_CUDA(cudaHostGetDevicePointer(&mapped, st, 0));
    d.ctl_src = mapped; d.ctl_dst = dev_ctl; d.ctl_n16 = int((total + 15) / 16); d.ctl_slot = dyn_slot;
    dyn_uploads_issued[dyn_slot]++;   // (the launch itself may be a replayed graph node)
  }
  else
  {
    SSR_CUDA(cudaMemcpyAsync(dev_ctl, st, total, cudaMemcpyHostToDevice, e->stream));
    SSR_CUDA(cudaEventRecord(dyn_copied[dyn_slot], e->stream));
  }
  dyn_tables.ctl = reinterpret_cast<const DynCtl*>(dev_ctl);
  dyn_tables.lists = reinterpret_cast<const int*>(dev_ctl + off_lists);
  dyn_tables.src_input = reinterpret_cast<const int*>(dev_ctl + off_input);
  dyn_tables.src_parts = reinterpret_cast<const int*>(dev_ctl + off_parts);
  dyn_tables.ring = dyn_ring.ptr;
  dyn_tables.upd = reinterpret_cast<const DynRingEntry*>(dev_ctl + off_upd);
  dyn_tables.stats = nullptr;
  dyn_tables.nsrc = N; dyn_tables.R = dyn_R; dyn_tables.Pmax = dyn_Pmax;
  dyn_wtab_dev = reinterpret_cast<const float*>(dev_ctl + off_w);
  dyn_upd_dev = reinterpret_cast<const DynRingEntry*>(dev_ctl + off_upd);
  prof.lap(1);

  // the inverse tables are static: one job per ear, one entry per accumulator kind
  if (inv_key != 1 || (inv.jobs.host.size() && inv.jobs.host[0].out != d_output))
  {
    inv.clear();
    const float scale = 1.0f / float(2 * B);
    for (int ear = 0; ear < 2; ++ear)
    {
      inv.jobs.host.push_back(InvJob{ear * 3, 3, d_output + size_t(ear) * B, level_ptr(N + ear)});
      for (int k = 0; k < 3; ++k) inv.entries.host.push_back({0, 0, 2, ear, k, scale});
    }
    inv.upload(e->stream);
    inv_key = 1;
  }
  prof.lap(2);

  build_fwd(d_input);
  prof.lap(3);
  d.N = N;
  d.kind_mask = 1 + 2 * dyn_slot;   // (graph key: the control buffers of the two parities have different addresses)
  // partitions read are counted by the kernel (Engine::d_dyn_stats); the
  // accumulators written per output are known here
  d.nrows = 2 * ((nk[0] > 0) + (nk[1] > 0) + (nk[2] > 0));
  d.graphable = true;
  return d;
}

// Prefilter bank: out[i] = in[i] * prefilter, every block
// (WfsRenderer::Input::Process, src/wfsrenderer.h:116-120: add_block; convolve()).
Renderer::StepDesc Renderer::prepare_bank(const float* d_input, float* d_output)
{
  InvPlan& inv = inv_plans[cur_slot];   // one set of inverse tables per output buffer
  int& inv_key = inv_keys[cur_slot];
  StepDesc d;
  d.d_input = d_input; d.d_output = d_output;
  const int N = int(sources.size());
  const int B = e->B;
  if (static_dirty)
  {
    static_mac.clear();
    static_mac.MT = 1;
    for (int n = 0; n < N; ++n)
    {
      static_mac.begin_group();
      for (int p = 0; p < partitions; ++p)
      {
        const float2* ptr = hrtfs->part_or_null(0, p);
        if (!ptr) continue;
        const float4* h = reinterpret_cast<const float4*>(ptr);
        static_mac.add_term(sources[n]->input->id, p, 0, &h);
      }
      static_mac.end_group();
    }
    static_mac.make_work(e->mac_slots(static_mac.MT), e->split_cost(static_mac.MT), e->mac_waves);
    static_mac.upload(e->stream);
    e->ensure_partials(size_t(static_mac.npartials) * static_mac.MT);
    static_dirty = false;
    invalidate_inv();
  }
  if (inv_key != 1 || (inv.jobs.host.size() && inv.jobs.host[0].out != d_output))
  {
    inv.clear();
    const float scale = 1.0f / float(2 * B);
    for (int n = 0; n < N; ++n)
    {
      const auto& grp = static_mac.groups[n];
      InvJob job{int(inv.entries.host.size()), 0, d_output + size_t(n) * B, level_ptr(N + n)};
      if (grp.term_end > grp.term_begin)
      {
        inv.entries.host.push_back({grp.partial_first, grp.nsplit, 1, 0, 0, scale});
        job.n_entries = 1;
      }
      inv.jobs.host.push_back(job);
    }
    inv.upload(e->stream);
    inv_key = 1;
  }
  const float one = 1.0f;
  build_fwd(d_input);
  e->upload_weights(&one, 1);
  const uint64_t parts = uint64_t(N) * uint64_t(hrtfs->valid[0]);
  d.N = N;
  d.kind_mask = 1;
  d.hparts[0] = parts; d.xparts[0] = parts; d.nrows = N;
  d.graphable = true;
  (void)B;
  return d;
}

// Multi-GPU: from now on the inverse kernel stores this rank's output blocks into
// the gather buffer (the root's own HBM, or the root's buffer through NVLink peer
// memory) and signals their arrival; d_output of step() is ignored.
void Renderer::bind_gather(Gather* g)
{
  if (outstanding) throw Error(SSR_ERR_STATE, "ssr_b200: bind_gather() while submitted blocks are pending");
  e->use_device();
  e->sync();
  if (g)
  {
    if (hostgather) throw Error(SSR_ERR_STATE, "ssr_b200: a host-direct gather is bound; unbind it first");
    if (cfg.kind != SSR_RENDERER_GENERIC)
      throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: only the Generic renderer shards its outputs (Binaural/BRS: replicas only)");
    if (g->e != e) throw Error(SSR_ERR_INVALID, "ssr_b200: gather belongs to another engine");
    if (g->M_total != M_total || g->counts[g->rank] != m_local || g->firsts[g->rank] != m_first)
      throw Error(SSR_ERR_INVALID, "ssr_b200: gather layout does not match the renderer's output shard");
    if (!g->ready()) throw Error(SSR_ERR_STATE, "ssr_b200: gather: buffer not connected (import the root's handle first)");
  }
  gather = g;
  invalidate_inv();
  drop_graphs();
  if (g && g->is_root()) h_out.resize(size_t(2) * M_total * e->B);
  else h_out.resize(size_t(2) * std::max(m_local, 1) * e->B);
}

// Multi-GPU, host-buffer call: the inverse kernel stores this rank's output blocks
// into the shared host segment (HostGather) through its device mapping.
void Renderer::bind_hostgather(HostGather* g)
{
  if (outstanding) throw Error(SSR_ERR_STATE, "ssr_b200: bind_hostgather() while submitted blocks are pending");
  e->use_device();
  e->sync();
  if (g)
  {
    if (gather) throw Error(SSR_ERR_STATE, "ssr_b200: a peer-memory gather is bound; unbind it first");
    if (cfg.kind != SSR_RENDERER_GENERIC)
      throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: only the Generic renderer shards its outputs (Binaural/BRS: replicas only)");
    if (g->e != e) throw Error(SSR_ERR_INVALID, "ssr_b200: hostgather belongs to another engine");
    if (g->M_total != M_total || g->counts[g->rank] != m_local || g->firsts[g->rank] != m_first)
      throw Error(SSR_ERR_INVALID, "ssr_b200: hostgather layout does not match the renderer's output shard");
  }
  hostgather = g;
  invalidate_inv();
  drop_graphs();
}

Renderer::StepDesc Renderer::prepare(const float* d_input, float* d_output)
{
  e->use_device();
  // slot 0 of the gather buffer; the kernel adds the slot offset of the block
  if (gather) d_output = gather->slot_ptr(0) + size_t(m_first) * e->B;
  if (hostgather) d_output = hostgather->d_slot_ptr(0) + size_t(m_first) * e->B;
  if (cfg.kind == SSR_RENDERER_GENERIC) return prepare_generic(d_input, d_output);
  if (cfg.kind == SSR_RENDERER_PREFILTER_BANK) return prepare_bank(d_input, d_output);
  return prepare_dynamic(d_input, d_output);
}

// Kernel launches of one block; nothing else (capturable into a CUDA graph).
void Renderer::launch(const StepDesc& d)
{
  const int B = e->B;
  InvPlan& inv = inv_plans[cur_slot];
  e->stage_begin();
  const bool dynamic = cfg.kind == SSR_RENDERER_BINAURAL || cfg.kind == SSR_RENDERER_BRS;
  // Generic plans on the bulk-copy MAC: the forward transform runs UNDER the MAC (only the
  // p == 0 terms need its result); the ring heads are then committed by the inverse kernel
  // (Binaural / BRS: the generated-terms MAC does the same)
  const bool overlap = e->fwd_overlap && d.N > 0 && !skip_forward
      && (dynamic || (cfg.kind == SSR_RENDERER_GENERIC && e->mac_tma() && static_mac.MT == kGenericMT));
  // block overlap (Engine::block_overlap): this block's forward transform and the p != 0 terms of its MAC may
  // run under the PREVIOUS block's inverse transform; the inverse kernel then commits the heads first
  // (Binaural / BRS: needs the control upload by kernel and the bulk-copy generated-terms MAC; the chain becomes
  //  forward -> control upload -> MAC, see ctl_upload_kernel.  Not when the forward kernel fetches its input from
  //  the caller's page-locked array (the synchronous host call): 256 KB of input reads on the PCIe link in front
  //  of the control data's 5 KB delay the MAC's start by more than the reordering gives -- measured, cfg3:
  //  55.5 -> 61.7 us per call; the upload-first chain keeps the link free for the control data there.)
  const bool block_overlap = overlap && e->block_overlap && e->pdl
      && (!dynamic || (d.ctl_n16 > 0 && e->dyn_tma() && dyn_block_overlap && !d.in_base));
  if (dynamic && block_overlap)
  {
    const FwdRingUpdate ru{dyn_tables.ctl, dyn_upd_dev, dyn_ring.ptr, dyn_R};
    e->launch_fwd(fwd->device(), d.N, nullptr, overlap, d.in_base, true, true);
    e->launch_ctl_upload(d.ctl_src, d.ctl_dst, d.ctl_n16, dyn_upload_count.ptr + d.ctl_slot, dyn_ack.ptr + d.ctl_slot,
        &ru, 2 * d.N);
  }
  else
  {
    if (d.ctl_n16 > 0)
    {
      e->launch_ctl_upload(d.ctl_src, d.ctl_dst, d.ctl_n16, dyn_upload_count.ptr + d.ctl_slot, dyn_ack.ptr + d.ctl_slot);
    }
    if (dynamic && d.N > 0)
    {
      // the forward kernel's CTA n also records source n's current filters in the ring
      const FwdRingUpdate ru{dyn_tables.ctl, dyn_upd_dev, dyn_ring.ptr, dyn_R};
      e->launch_fwd(fwd->device(), d.N, &ru, overlap, d.in_base);
    }
    else if (!skip_forward) e->launch_fwd(fwd->device(), d.N, nullptr, overlap, d.in_base, block_overlap);
  }
  e->stage_mark(1);
  if (cfg.kind == SSR_RENDERER_GENERIC)
  {
    for (int k = 0; k < 3; ++k)
    {
      if (!(d.kind_mask & (1 << k))) continue;
      e->launch_mac(static_mac, k * d.N, k * static_mac.npartials, nullptr, overlap);
      if (e->cur)
      {
        e->cur->mac_bytes += uint64_t(8) * B * (d.hparts[k] + d.xparts[k] + uint64_t(d.nrows));
        e->cur->mac_terms += d.hparts[k];
      }
    }
  }
  else
  {
    if (cfg.kind == SSR_RENDERER_PREFILTER_BANK) e->launch_mac(static_mac, 0, 0);
    else
    {
      e->launch_mac_dyn(dyn_tables, dyn_G, dyn_wtab_dev, overlap);
    }
    if (e->cur)
    {
      e->cur->mac_bytes += uint64_t(8) * B * (d.hparts[0] + d.xparts[0] + uint64_t(d.nrows));
      e->cur->mac_terms += d.hparts[0];
    }
  }
  e->stage_mark(2);
  const HeadCommit hc{overlap ? fwd->device() : nullptr, overlap ? d.N : 0,
                      block_overlap ? 2 : ((e->inv_trigger && e->pdl) ? 1 : 0)};
  if (cfg.kind == SSR_RENDERER_BINAURAL || cfg.kind == SSR_RENDERER_BRS)
    e->launch_inv_dyn(inv, dyn_tables.ctl, dyn_Pmax, dyn_G, hc, ctl_by_kernel && e->pdl);
  else if (gather) e->launch_inv(inv, gather->next_block(), hc);
  else if (hostgather)
  {
    // the slot this block goes to must have been read out on the root (host-side wait;
    // forward transform and MAC are already queued and run meanwhile)
    hostgather->wait_slot_free(hostgather->epoch);
    e->launch_inv(inv, hostgather->next_block(int(inv.jobs.uploaded)), hc);
  }
  else e->launch_inv(inv, GatherArgs{}, hc);
  e->stage_mark(3);
  e->stage_end();
}

void Renderer::step(const float* d_input, float* d_output)
{
  if (hostgather)
    throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: the device-buffer step delivers through the peer-memory gather "
        "(ssr_gather_*); a host-direct gather serves ssr_renderer_process");
  apply_controls();
  cur_slot = 0;
  const StepDesc d = prepare(d_input, d_output);
  prof.start();
  launch(d);
  prof.lap(4);
  prof.n++;
}

void Renderer::drop_graphs()
{
  for (auto& per_slot : graphs)
    for (auto& g : per_slot)
    {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g = GraphEntry();
    }
}

// true if the B-float channel blocks behind `ptrs` form one contiguous array in
// page-locked host memory (then the DMA engine can use them in place)
static bool contiguous_pinned(Engine* e, const float* const* ptrs, int count, int B)
{
  if (count <= 0 || !ptrs || !ptrs[0]) return false;
  for (int i = 1; i < count; ++i)
    if (ptrs[i] != ptrs[0] + size_t(i) * B) return false;
  return e->is_pinned(ptrs[0], size_t(count) * B * sizeof(float));
}

// ------------------------------------------------------------------ offline, time-batched
// Would this block leave every source's weight as it is (and non-fading)?  Then
// kTimeBatch such blocks can share one pass over the filter spectra.
bool Renderer::batch_eligible()
{
  if (cfg.kind != SSR_RENDERER_GENERIC || gather || hostgather || static_dirty || sources.empty()) return false;
  if (!e->mac_tma() || static_mac.MT != kGenericMT || e->timing || metering) return false;
  for (auto& sp : sources)
  {
    Source& s = *sp;
    const float next = (!processing || s.mute) ? 0.0f : s.gain * master_volume * mvc;  // base_weight()
    if (next != s.weighting_factor.get() || next != s.w.get() || s.w.changed()) return false;
  }
  return true;
}

void Renderer::run_batch(const float* const* in, float* const* out)
{
  constexpr int T = Input::kTimeBatch;
  e->use_device();
  const int N = int(sources.size());
  const int B = e->B;
  const size_t in_floats = size_t(T) * N * B, out_floats = size_t(T) * m_local * B;
  if (d_in_batch.count < in_floats || d_out_batch.count < out_floats)
  {
    e->sync();
    d_in_batch.resize(in_floats); h_in_batch.resize(in_floats);
    d_out_batch.resize(out_floats); h_out_batch.resize(out_floats);
    fwd_cache.clear();
    inv_batch_key = -1;
  }
  // page-locked arrays laid out [t][channel][B] are used in place, like process() does
  const float* hin = h_in_batch.ptr;
  if (contiguous_pinned(e, in, T * N, B)) hin = in[0];
  else
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < N; ++i)
        std::memcpy(h_in_batch.ptr + (size_t(t) * N + i) * B, in[size_t(t) * N + i], B * sizeof(float));
  float* hout = h_out_batch.ptr;
  const bool out_in_place = contiguous_pinned(e, out, T * m_local, B);
  if (out_in_place) hout = out[0];
  SSR_CUDA(cudaMemcpyAsync(d_in_batch.ptr, hin, in_floats * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  e->stats.h2d_bytes += in_floats * sizeof(float);
  e->stats.d2h_bytes += out_floats * sizeof(float);

  // the per-block state updates of T constant blocks (src/genericrenderer.h:122-129)
  wtab_host.assign(size_t(3) * N, 0.0f);
  for (int t = 0; t < T; ++t)
    for (int n = 0; n < N; ++n)
    {
      Source& s = *sources[n];
      base_weight(s);
      s.w.assign(s.weighting_factor.get());
      s.mode = mode_from_weight(s.w);
      wtab_host[n] = s.w.get();   // constant accumulator kind; zero weights are dropped by the kernel
    }
  e->upload_weights(wtab_host.data(), wtab_host.size());

  // T forward transforms (each advances the delay line by one block)
  for (int t = 0; t < T; ++t)
  {
    build_fwd(d_in_batch.ptr + size_t(t) * N * B);
    e->launch_fwd(fwd->device(), N);
  }
  // one pass over the filter spectra for all T blocks
  e->ensure_partials(size_t(static_mac.npartials) * T * static_mac.MT);
  e->launch_mac_batch(static_mac, 0);
  // T x outputs inverse transforms; partial layout [CTA][t][row]
  const int key = static_mac.npartials * 131 + m_local;
  if (inv_batch_key != key || (inv_batch.jobs.host.size() && inv_batch.jobs.host[0].out != d_out_batch.ptr))
  {
    inv_batch.clear();
    const float scale = 1.0f / float(2 * B);
    for (int t = 0; t < T; ++t)
      for (int m = 0; m < m_local; ++m)
      {
        const int g = m / kGenericMT, row = m % kGenericMT;
        const auto& grp = static_mac.groups[g];
        inv_batch.entries.host.push_back({grp.partial_first, grp.nsplit, T * kGenericMT, t * kGenericMT + row, 0, scale});
        inv_batch.jobs.host.push_back(InvJob{int(inv_batch.entries.host.size()) - 1, 1,
            d_out_batch.ptr + (size_t(t) * m_local + m) * B, nullptr});
      }
    inv_batch.upload(e->stream);
    inv_batch_key = key;
  }
  e->launch_inv(inv_batch);
  SSR_CUDA(cudaMemcpyAsync(hout, d_out_batch.ptr, out_floats * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  e->sync();
  if (!out_in_place)
    for (int t = 0; t < T; ++t)
      for (int m = 0; m < m_local; ++m)
        std::memcpy(out[size_t(t) * m_local + m], h_out_batch.ptr + (size_t(t) * m_local + m) * B, B * sizeof(float));
  batched_blocks += T;
}

void Renderer::process_batch(int n, int nblocks, const float* const* in, float* const* out)
{
  if (outstanding) throw Error(SSR_ERR_STATE, "ssr_b200: process_batch() while submitted blocks are pending");
  if (n != e->B) throw Error(SSR_ERR_INVALID, "ssr_b200: audio_callback: n != block size");
  if (nblocks < 0) throw Error(SSR_ERR_INVALID, "ssr_b200: process_batch: negative block count");
  const int N = int(sources.size());
  const int n_out = host_outputs();
  int t = 0;
  while (t < nblocks)
  {
    // the commands queued so far take effect with this block (and then with none of
    // the following ones of a batched run: a control thread that wants per-block control
    // of an offline render calls this function block by block)
    apply_controls();
    if (nblocks - t >= Input::kTimeBatch && batch_eligible())
    {
      run_batch(in + size_t(t) * N, out + size_t(t) * n_out);
      blocks_started += Input::kTimeBatch - 1;
      t += Input::kTimeBatch;
    }
    else
    {
      controls_applied = true;
      process(n, in + size_t(t) * N, out + size_t(t) * n_out);
      t += 1;
    }
  }
}

void Renderer::invalidate_inv()
{
  for (int i = 0; i < 2; ++i) { inv_keys[i] = -1; inv_plans[i].clear(); }
}

void Renderer::process(int n, const float* const* in, float* const* out)
{
  if (outstanding) throw Error(SSR_ERR_STATE, "ssr_b200: process() while submitted blocks are pending");
  // synchronous call: the caller's buffers stay valid until we return, so pinned
  // contiguous channel arrays are used in place (no staging copies); everything
  // runs on the engine's stream (lowest latency for one block)
  submit_block(n, in, out, false);
  wait(out);
}

// Pipelined: copy-in stream -> engine stream -> copy-out stream, two blocks in
// flight, so block t+1's H2D and block t-1's D2H overlap block t's kernels
// (north_star item 5: the reference's per-block thread fan-out,
// apf/apf/mimoprocessor.h:452-483, becomes a CUDA-stream pipeline).
void Renderer::submit(int n, const float* const* in) { submit_block(n, in, nullptr, true); }

void Renderer::submit_block(int n, const float* const* in, float* const* direct_out, bool pipelined)
{
  // assert(n == block_size()) (apf/apf/pointer_policy.h:114)
  if (n != e->B) throw Error(SSR_ERR_INVALID, "ssr_b200: audio_callback: n != block size");
  if (outstanding >= 2) throw Error(SSR_ERR_STATE, "ssr_b200: at most 2 blocks may be in flight");
  if (!controls_applied) apply_controls();
  controls_applied = false;
  e->use_device();
  const int N = int(sources.size());
  const int B = e->B;
  const size_t in_floats = size_t(std::max(N, 1)) * B;
  if (h_in.count < 2 * in_floats)
  {
    e->sync();
    h_in.resize(2 * in_floats);
    for (auto& d : d_in) d.resize(in_floats);
    fwd_cache.clear();
  }
  static_assert(sizeof(float) == 4, "");
  const int slot = int(submitted & 1);
  cur_slot = pipelined ? slot : 0;   // device buffers / inverse tables / graph of this block
  if (pipelined && !s_in)
  {
    SSR_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    SSR_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i)
    {
      SSR_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
      SSR_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
    }
  }
  float* din = d_in[cur_slot].ptr;
  float* dout = d_out[cur_slot].ptr;
  // with a gather bound the root delivers ALL outputs, the other ranks none
  const int n_host_out = host_outputs();
  if (hostgather && pipelined) throw Error(SSR_ERR_UNSUPPORTED, "ssr_b200: submit()/wait() with a host-direct gather: use process()");
  if (hostgather && hostgather->is_root()) hostgather->publish_consumed(hostgather->epoch);
  const float* hin = nullptr;
  if (!pipelined && N > 0 && contiguous_pinned(e, in, N, B)) hin = in[0];  // synchronous call: used in place
  else
  {
    float* stage = h_in.ptr + size_t(slot) * in_floats;
    for (int i = 0; i < N; ++i) std::memcpy(stage + size_t(i) * B, in[i], B * sizeof(float));
    hin = stage;
  }
  float* hout = hostgather ? nullptr : h_out.ptr + size_t(slot) * n_host_out * B;
  slot_direct[slot] = false;
  if (hostgather)
  {
    // the caller reads the blocks in place when `out` points at the segment's slot of this block
    const float* want = hostgather->slot_ptr(int(hostgather->epoch & 1));
    bool in_place = direct_out && n_host_out > 0;
    for (int m = 0; in_place && m < n_host_out; ++m) in_place = direct_out[m] == want + size_t(m) * B;
    slot_direct[slot] = in_place;
  }
  else if (direct_out && n_host_out > 0 && contiguous_pinned(e, direct_out, n_host_out, B))
  {
    hout = direct_out[0];
    slot_direct[slot] = true;
  }
  e->stats.h2d_bytes += uint64_t(N) * B * sizeof(float);
  e->stats.d2h_bytes += uint64_t(n_host_out) * B * sizeof(float);
  if (!done[slot]) SSR_CUDA(cudaEventCreateWithFlags(&done[slot], cudaEventDisableTiming));
  // Synchronous call into the caller's page-locked output array: the inverse kernel
  // stores its blocks straight into it (mapped host memory, posted PCIe writes that
  // overlap the kernel) instead of a device buffer + D2H copy afterwards.
  bool zero_copy_out = zero_copy && !pipelined && !gather && !hostgather && slot_direct[slot];
  if (zero_copy_out)
  {
    void* mapped = nullptr;  // the device's view of the array (the same address under UVA)
    if (cudaHostGetDevicePointer(&mapped, hout, 0) == cudaSuccess && mapped) dout = static_cast<float*>(mapped);
    else { cudaGetLastError(); zero_copy_out = false; }
  }

  StepDesc d = prepare(din, dout);
  // Synchronous call with the caller's page-locked input array: the forward kernel reads the
  // blocks straight from it (mapped host memory) -- no H2D copy in front of the block, and
  // with the forward transform running under the MAC the PCIe transfer is off the critical
  // path as well.  (Every sample still crosses the bus once: the kernel keeps the block as
  // the next block's first half in device memory.)
  bool direct_in = zero_copy_in && !pipelined && N > 0 && hin == in[0] && e->pdl
      && !(e->fft1024 && N >= Engine::kFft1024MinJobs);
  if (direct_in)
  {
    void* mapped = nullptr;
    if (cudaHostGetDevicePointer(&mapped, const_cast<float*>(hin), 0) == cudaSuccess && mapped)
    {
      d.in_base = static_cast<const float*>(mapped);
      d.graphable = false;   // the array's address is a kernel argument of this block
    }
    else { cudaGetLastError(); direct_in = false; }
  }
  if (pipelined)
  {
    // H2D on the copy-in stream, once the kernels of block t-2 (readers of this
    // input buffer) are through; the engine stream then waits for the copy and for
    // block t-2's D2H out of this output buffer
    if (slot_used[slot]) SSR_CUDA(cudaStreamWaitEvent(s_in, ev_k[slot], 0));
    if (N > 0)
      SSR_CUDA(cudaMemcpyAsync(din, hin, size_t(N) * B * sizeof(float), cudaMemcpyHostToDevice, s_in));
    SSR_CUDA(cudaEventRecord(ev_in[slot], s_in));
    SSR_CUDA(cudaStreamWaitEvent(e->stream, ev_in[slot], 0));
    if (slot_used[slot]) SSR_CUDA(cudaStreamWaitEvent(e->stream, done[slot], 0));
  }
  else if (N > 0 && !direct_in)
    SSR_CUDA(cudaMemcpyAsync(din, hin, size_t(N) * B * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  // Steady state (no table moved or resized since the last blocks): the kernel
  // sequence {forward FFT, MAC passes, [partial reduce], inverse FFT} is replayed
  // as one CUDA graph per device-buffer slot and accumulator-kind mask; the two
  // copies stay outside so that they can use the caller's buffers.  Per-stage
  // event timing needs separate launches.
  bool done_by_graph = false;
  if (use_graphs && d.graphable && !e->timing && d.kind_mask > 0 && d.kind_mask < 8)
  {
    GraphEntry& g = graphs[cur_slot][d.kind_mask];
    const uint64_t epoch = e->graph_epoch();
    if (g.exec && g.epoch == epoch)
    {
      SSR_CUDA(cudaGraphLaunch(g.exec, e->stream));
      ++graph_launches;
      e->count_launch(g.kernels);
      done_by_graph = true;
    }
    else if (g.epoch == epoch && ++g.seen >= 2)
    {
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
      cudaGraph_t graph = nullptr;
      const uint64_t before = e->stats.launches_total;
      SSR_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
      try { launch(d); }
      catch (...) { cudaStreamEndCapture(e->stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
      SSR_CUDA(cudaStreamEndCapture(e->stream, &graph));
      g.kernels = int(e->stats.launches_total - before);  // counted once here, once per replay below
      cudaError_t err = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (err == cudaSuccess)
      {
        SSR_CUDA(cudaGraphLaunch(g.exec, e->stream));
        ++graph_launches;
        done_by_graph = true;
      }
      else { g.exec = nullptr; use_graphs = false; }  // fall back to plain launches for good
    }
    else if (g.epoch != epoch) { g.epoch = epoch; g.seen = 0; if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; } }
  }
  if (!done_by_graph) launch(d);

  const float* src = dout;
  size_t out_floats = size_t(m_local) * B;
  if (hostgather)
  {
    out_floats = 0;   // the inverse kernel delivered the blocks; no copy on any rank
    slot_epoch[slot] = hostgather->epoch - 1;
  }
  if (gather)
  {
    out_floats = 0;
    if (gather->is_root())
    {
      src = gather->wait();  // every rank's blocks of this block have arrived (engine stream)
      out_floats = size_t(M_total) * B;
    }
  }
  cudaStream_t copy_stream = e->stream;
  if (pipelined)
  {
    SSR_CUDA(cudaEventRecord(ev_k[slot], e->stream));
    SSR_CUDA(cudaStreamWaitEvent(s_out, ev_k[slot], 0));
    copy_stream = s_out;
  }
  if (out_floats && !zero_copy_out)
    SSR_CUDA(cudaMemcpyAsync(hout, src, out_floats * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
  if (gather && gather->is_root()) gather->release(copy_stream);   // after the slot has been read out
  SSR_CUDA(cudaEventRecord(done[slot], copy_stream));
  slot_used[slot] = pipelined;
  submitted++;
  outstanding++;
}

void Renderer::wait(float* const* out)
{
  if (!outstanding) throw Error(SSR_ERR_STATE, "ssr_b200: wait() without a submitted block");
  e->use_device();
  const int B = e->B;
  const int slot = int((submitted - outstanding) & 1);
  SSR_CUDA(cudaEventSynchronize(done[slot]));
  const int n_host_out = host_outputs();
  if (gather && gather->timed_out())
  {
    // a wait on another rank gave up (inverse.cuh, gather_wait_ge): the slot holds a
    // stale or partial block -- fail the call instead of delivering it
    outstanding--;
    throw Error(SSR_ERR_STATE, "ssr_b200: gather: a wait on a peer rank timed out; the output block is not valid");
  }
  if (hostgather && hostgather->is_root())
  {
    // our own kernels are through; now every other rank's flag (host memory)
    try { hostgather->wait_all(slot_epoch[slot]); }
    catch (...) { outstanding--; throw; }
  }
  if (!slot_direct[slot])
  {
    const float* hout = hostgather ? hostgather->slot_ptr(int(slot_epoch[slot] & 1))
                                   : h_out.ptr + size_t(slot) * n_host_out * B;
    for (int m = 0; m < n_host_out; ++m) std::memcpy(out[m], hout + size_t(m) * B, B * sizeof(float));
  }
  outstanding--;
}

}  // namespace ssr_b200
