// nonuniform.cu -- non-uniformly partitioned Generic renderer (NOT in the reference:
// doc/manual/renderers.tex:199-204 only notes the latency / cost trade-off of the partition size).
//
// The reference convolves every (source, output) pair with P partitions of the block size B, so each
// block streams the WHOLE filter set once: 8 B bytes x N x M x P.  Here the head of every impulse
// response keeps partitions of B (level 0, exactly the reference's algorithm: zero added latency, the
// reference's per-block weight crossfade), and the tail is cut into partitions of K B samples
// (levels 1..), which are due only every K blocks: a level-l spectrum is read once per K blocks
// instead of once per block, and the bytes per block fall from P to P0 + sum_l P_l partitions'
// worth (cfg5: 47 -> 18 with two levels 8 x 1024 + 10 x 4096).
//
// Every level is the SAME engine at a larger block size: an ssr_b200 Engine with block K B and
// Generic renderers over the IR slice [tap_first, tap_first + P_l K B).  Scheduling is what makes it a
// real-time algorithm (block t's output leaves before block t + 1's input exists) with an even load:
//   * the K B-sample input segment i of a level is complete after block (i + 1) K - 1;
//   * the level's outputs are split over K sub-renderers; sub-renderer j transforms, multiplies and
//     inverse-transforms ITS outputs for segment i in block (i + 1) K + j (time-distributed: every block
//     carries 1 / K of the level's work);
//   * the result segment belongs to output times [i K B + tap_first, ...), i.e. it is first needed in
//     block i K + tap_first / B >= (i + 2) K: hence tap_first >= 2 K B (and a multiple of K B);
//   * nu_mix_kernel adds the due B-sample slice of every level's result ring to level 0's block.
// Source weights: level 0 applies the reference's per-block crossfade; a level-l sub-renderer applies
// the same rule at ITS block rate (a gain change reaches the late tail up to one long partition later
// and fades over K B samples).  With constant weights the output equals the reference's within
// rounding (tests/test_gpu_nonuniform.py: <= 1e-5 of full scale against float64 direct convolution).
#include "engine.hpp"

#include <algorithm>

namespace ssr_b200
{

constexpr int kNuMaxLevels = 4;

struct NuMixArgs
{
  const float* src[kNuMaxLevels];   // level result slice of this block: row m at src[l] + m * stride[l]
  long stride[kNuMaxLevels];
  int n;
};

// out[m][0..B) += sum_l src_l[m][0..B); grid = rows, B % 4 == 0
__global__ void __launch_bounds__(256)
nu_mix_kernel(float* __restrict__ out, const NuMixArgs a, int B)
{
  float4* o = reinterpret_cast<float4*>(out + size_t(blockIdx.x) * B);
  for (int i = threadIdx.x; i < B / 4; i += blockDim.x)
  {
    float4 v = o[i];
    for (int l = 0; l < a.n; ++l)
    {
      const float4 s = __ldcs(reinterpret_cast<const float4*>(a.src[l] + size_t(blockIdx.x) * a.stride[l]) + i);
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    o[i] = v;
  }
}

struct NuRenderer
{
  struct Level
  {
    int K = 1;                 // block multiple
    int parts = 0;             // partitions of K B taps
    long tap_first = 0;        // IR slice of the level: [tap_first, tap_first + parts K B)
    int D = 0;                 // tap_first / (K B): result segment i is consumed in period i + D
    std::unique_ptr<Engine> eng;
    std::vector<std::unique_ptr<Renderer>> subs;   // level 0: one; level l: <= K, over disjoint output ranges
    std::vector<int> sub_first;                    // first local output row of each sub-renderer
    DeviceArray<float> in_acc[2];                  // [sources][K B]: segment being filled / being rendered
    DeviceArray<float> out_ring;                   // [D][m_local][K B]
  };

  ssr_nu_config cfg;
  int B, M_total, m_first, m_local;
  cudaStream_t stream = nullptr;
  std::vector<std::unique_ptr<Level>> levels;
  int n_sources = 0;
  uint64_t block = 0;          // blocks rendered so far
  uint64_t mix_launches = 0;
  bool share_inputs = std::getenv("SSR_B200_NU_SHARE_INPUTS") == nullptr || std::atoi(std::getenv("SSR_B200_NU_SHARE_INPUTS")) != 0;
  DeviceArray<float> d_in, d_out;
  PinnedArray<float> h_in, h_out;
  std::vector<int> ids;        // source ids (the same in every sub-renderer)

  explicit NuRenderer(const ssr_nu_config& c) : cfg(c)
  {
    B = c.block_size;
    if (c.levels < 1 || c.levels > kNuMaxLevels) throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: 1..4 levels");
    if (c.outputs <= 0) throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: outputs > 0");
    M_total = c.outputs;
    m_first = c.output_first;
    m_local = c.output_count > 0 ? c.output_count : M_total - m_first;
    if (m_first < 0 || m_local <= 0 || m_first + m_local > M_total) throw Error(SSR_ERR_INVALID, "ssr_b200: bad output shard");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || c.device < 0 || c.device >= ndev)
      throw Error(SSR_ERR_CUDA, "ssr_b200: no CUDA device usable; this engine has no CPU fallback");
    SSR_CUDA(cudaSetDevice(c.device));
    SSR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    long tap = 0;
    for (int l = 0; l < c.levels; ++l)
    {
      std::unique_ptr<Level> Lp(new Level);
      Level& L = *Lp;
      L.K = c.level_mult[l];
      L.parts = c.level_partitions[l];
      if (l == 0 && L.K != 1) throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: level 0 has the block size itself (level_mult[0] = 1)");
      if (L.K < 1 || (L.K & (L.K - 1)) != 0 || (l > 0 && L.K <= levels[size_t(l) - 1]->K))
        throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: level_mult must be increasing powers of two");
      if (L.parts < 1) throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: every level needs >= 1 partition");
      const long BL = long(L.K) * B;
      L.tap_first = tap;
      if (l > 0 && (tap % BL != 0 || tap < 2 * BL))
        throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: a level must start at a multiple of its partition size, "
            "at least two of its partitions into the impulse response (time-distributed schedule)");
      L.D = int(tap / BL);
      tap += long(L.parts) * BL;
      ssr_engine_config ec{};
      ec.device = c.device; ec.block_size = int(BL); ec.sample_rate = c.sample_rate; ec.stream = stream;
      L.eng.reset(new Engine(ec));
      // sub-renderers: K output ranges (fewer when there are fewer outputs than K)
      const int nsub = l == 0 ? 1 : std::min(L.K, m_local);
      for (int j = 0; j < nsub; ++j)
      {
        const int lo = int(long(m_local) * j / nsub), hi = int(long(m_local) * (j + 1) / nsub);
        ssr_renderer_config rc{};
        rc.kind = SSR_RENDERER_GENERIC;
        rc.outputs = M_total; rc.output_first = m_first + lo; rc.output_count = hi - lo;
        rc.master_volume_correction_db = c.master_volume_correction_db;
        L.subs.emplace_back(new Renderer(L.eng.get(), rc));
        L.subs.back()->tail_level = l > 0;   // no initial fade-in: by the time a tail level sounds, the reference's is over
        L.sub_first.push_back(lo);
      }
      if (l > 0)
      {
        L.out_ring.resize(size_t(L.D) * m_local * BL);
        SSR_CUDA(cudaMemsetAsync(L.out_ring.ptr, 0, L.out_ring.count * sizeof(float), stream));
      }
      levels.push_back(std::move(Lp));
    }
    d_out.resize(size_t(m_local) * B);
    h_out.resize(size_t(m_local) * B);
    SSR_CUDA(cudaStreamSynchronize(stream));
  }

  ~NuRenderer()
  {
    cudaSetDevice(cfg.device);
    if (stream) cudaStreamSynchronize(stream);
    levels.clear();   // engines (and their use of the stream) go first
    if (stream) cudaStreamDestroy(stream);
  }

  long taps_covered() const { const Level& L = *levels.back(); return L.tap_first + long(L.parts) * L.K * B; }

  void grow_inputs()
  {
    SSR_CUDA(cudaSetDevice(cfg.device));
    d_in.resize(size_t(n_sources) * B);
    h_in.resize(size_t(n_sources) * B);
    for (size_t l = 1; l < levels.size(); ++l)
      for (auto& a : levels[l]->in_acc)
      {
        a.resize(size_t(n_sources) * levels[l]->K * B);
        SSR_CUDA(cudaMemsetAsync(a.ptr, 0, a.count * sizeof(float), stream));
      }
  }

  // ir: frames x M_total interleaved (as ssr_renderer_add_source); or synthetic (ir == null)
  int add_source(const float* ir, long frames, int channels, bool synthetic, uint64_t seed, uint64_t id_offset,
                 const float* envelope, float scale)
  {
    if (block != 0) throw Error(SSR_ERR_STATE, "ssr_b200: nonuniform: sources are added before the first block");
    if (channels != M_total)
      throw Error(SSR_ERR_FILE, "apf::load_sndfile(): IR set has " + std::to_string(channels)
          + " channels instead of " + std::to_string(M_total) + "!");
    if (frames <= 0) throw Error(SSR_ERR_INVALID, "ssr_b200: add_source: empty IR set");
    if (frames > taps_covered())
      throw Error(SSR_ERR_INVALID, "ssr_b200: nonuniform: impulse response longer than the configured levels cover ("
          + std::to_string(taps_covered()) + " taps)");
    int sid = -1;
    for (auto& Lp : levels)
    {
      Level& L = *Lp;
      const long BL = long(L.K) * B;
      // this level's slice of the IR; sources whose IR ends before the slice get one silent partition
      const long avail = std::max<long>(0, std::min<long>(frames - L.tap_first, long(L.parts) * BL));
      for (auto& rp : L.subs)
      {
        Renderer& R = *rp;
        const int P = int(std::max<long>(1, (avail + BL - 1) / BL));
        std::unique_ptr<FilterSet> fs(new FilterSet(R.e, R.m_local, P));
        if (avail > 0)
        {
          if (synthetic)
            fs->generate(0, R.m_local, avail, seed, id_offset + uint64_t(R.m_first), envelope ? envelope + L.tap_first : nullptr,
                scale, uint64_t(L.tap_first));
          else
            fs->load_time(0, R.m_local, ir + size_t(L.tap_first) * channels + R.m_first, avail, channels, 1);
        }
        const int id = R.add_source(std::move(fs), channels);
        if (sid < 0) sid = id;
        else if (id != sid) throw Error(SSR_ERR_STATE, "ssr_b200: nonuniform: source ids of the levels diverged");
        // ONE delay line per (level, source): sub-renderer 0 transforms the level's input segment in its phase,
        // the others (later phases of the same period) multiply against the same spectra
        if (share_inputs && &R != L.subs[0].get())
        {
          if (R.sources.back()->input->P != L.subs[0]->sources.back()->input->P)
            throw Error(SSR_ERR_STATE, "ssr_b200: nonuniform: partition counts of a level's sub-renderers differ");
          R.sources.back()->input = L.subs[0]->sources.back()->input;
          R.skip_forward = true;
        }
      }
    }
    ids.push_back(sid);
    n_sources++;
    grow_inputs();
    return sid;
  }

  template<typename F> void for_all(F f) { for (auto& L : levels) for (auto& r : L->subs) f(*r); }

  void push(int kind, int id, float a)
  {
    for_all([&](Renderer& r) {
      if (kind == ControlQueue::SRC_GAIN || kind == ControlQueue::SRC_MUTE) (void)r.find(id);
      if (!r.control.push(ControlQueue::Cmd{kind, id, a, 0.0f}))
        throw Error(SSR_ERR_STATE, "ssr_b200: control queue full");
    });
  }

  // one block, device buffers: d_input [sources][B], d_output [local outputs][B]; async on `stream`
  void step(const float* d_input, float* d_output)
  {
    SSR_CUDA(cudaSetDevice(cfg.device));
    const uint64_t n = block;
    levels[0]->subs[0]->step(d_input, d_output);
    NuMixArgs mix{};
    for (size_t l = 1; l < levels.size(); ++l)
    {
      Level& L = *levels[l];
      const long BL = long(L.K) * B;
      const uint64_t period = n / uint64_t(L.K);
      const int phase = int(n % uint64_t(L.K));
      // this block's samples join the segment that is being filled
      if (n_sources > 0)
        SSR_CUDA(cudaMemcpy2DAsync(L.in_acc[period & 1].ptr + size_t(phase) * B, size_t(BL) * sizeof(float),
              d_input, size_t(B) * sizeof(float), size_t(B) * sizeof(float), size_t(n_sources),
              cudaMemcpyDeviceToDevice, stream));
      // sub-renderer `phase` renders the segment completed at the end of the previous period
      if (period >= 1 && phase < int(L.subs.size()))
      {
        const uint64_t seg = period - 1;
        float* dst = L.out_ring.ptr + (size_t(seg % uint64_t(L.D)) * m_local + size_t(L.sub_first[size_t(phase)])) * BL;
        L.subs[size_t(phase)]->step(L.in_acc[seg & 1].ptr, dst);
      }
      // the slice of the level's results due in this block: segment period - D
      if (period >= uint64_t(L.D))
      {
        const uint64_t seg = period - uint64_t(L.D);
        mix.src[mix.n] = L.out_ring.ptr + size_t(seg % uint64_t(L.D)) * m_local * BL + size_t(phase) * B;
        mix.stride[mix.n] = BL;
        mix.n++;
      }
    }
    if (mix.n > 0)
    {
      nu_mix_kernel<<<m_local, 256, 0, stream>>>(d_output, mix, B);
      SSR_CUDA(cudaGetLastError());
      mix_launches++;
    }
    block++;
  }

  // = pointer_policy::audio_callback(n, in, out) (apf/apf/pointer_policy.h:112-122): n frames = the block size
  void process(int n, const float* const* in, float* const* out)
  {
    if (n != B) throw Error(SSR_ERR_INVALID, "ssr_b200: process(): n must equal the block size");
    const int nin = n_sources;
    SSR_CUDA(cudaSetDevice(cfg.device));
    // channel blocks that form one contiguous page-locked array are used in place by the copy engines
    // (the layout a driver with its own pinned I/O buffers passes); anything else is staged
    Engine* e0 = levels[0]->eng.get();
    auto contiguous_pinned = [&](const float* const* ptrs, int count) {
      if (count <= 0 || !ptrs || !ptrs[0]) return false;
      for (int i = 1; i < count; ++i) if (ptrs[i] != ptrs[0] + size_t(i) * B) return false;
      return e0->is_pinned(ptrs[0], size_t(count) * B * sizeof(float));
    };
    const float* src = h_in.ptr;
    if (contiguous_pinned(in, nin)) src = in[0];
    else for (int i = 0; i < nin; ++i) std::memcpy(h_in.ptr + size_t(i) * B, in[i], size_t(B) * sizeof(float));
    if (nin > 0)
      SSR_CUDA(cudaMemcpyAsync(d_in.ptr, src, size_t(nin) * B * sizeof(float), cudaMemcpyHostToDevice, stream));
    step(d_in.ptr, d_out.ptr);
    const bool direct_out = contiguous_pinned(out, m_local);
    float* dst = direct_out ? out[0] : h_out.ptr;
    SSR_CUDA(cudaMemcpyAsync(dst, d_out.ptr, size_t(m_local) * B * sizeof(float), cudaMemcpyDeviceToHost, stream));
    SSR_CUDA(cudaStreamSynchronize(stream));
    if (!direct_out)
      for (int m = 0; m < m_local; ++m) std::memcpy(out[m], h_out.ptr + size_t(m) * B, size_t(B) * sizeof(float));
  }
};

}  // namespace ssr_b200

using namespace ssr_b200;

struct ssr_nu { NuRenderer impl; explicit ssr_nu(const ssr_nu_config& c) : impl(c) {} };

// (unity build: guarded() and the ssr_last_error() slot of capi.cu are in scope)

extern "C"
{

int ssr_nu_create(const ssr_nu_config* cfg, ssr_nu** out)
{
  return guarded([&] {
    if (!cfg || !out) throw Error(SSR_ERR_INVALID, "ssr_b200: nu_create: null argument");
    *out = nullptr;
    *out = new ssr_nu(*cfg);
  });
}

void ssr_nu_destroy(ssr_nu* r) { delete r; }

int ssr_nu_add_source(ssr_nu* r, const float* ir, long frames, int channels, int* id)
{
  return guarded([&] {
    if (!r || !ir) throw Error(SSR_ERR_INVALID, "ssr_b200: nu_add_source: null argument");
    const int sid = r->impl.add_source(ir, frames, channels, false, 0, 0, nullptr, 1.0f);
    if (id) *id = sid;
  });
}

int ssr_nu_add_source_synthetic(ssr_nu* r, long taps, int channels, uint64_t seed, uint64_t channel_id_offset,
                                const float* envelope, float scale, int* id)
{
  return guarded([&] {
    if (!r) throw Error(SSR_ERR_INVALID, "ssr_b200: null renderer");
    const int sid = r->impl.add_source(nullptr, taps, channels, true, seed, channel_id_offset, envelope, scale);
    if (id) *id = sid;
  });
}

int ssr_nu_source_set_gain(ssr_nu* r, int id, float gain)
{
  return guarded([&] { if (!r) throw Error(SSR_ERR_INVALID, "ssr_b200: null renderer"); r->impl.push(ControlQueue::SRC_GAIN, id, gain); });
}

int ssr_nu_source_set_mute(ssr_nu* r, int id, int mute)
{
  return guarded([&] { if (!r) throw Error(SSR_ERR_INVALID, "ssr_b200: null renderer"); r->impl.push(ControlQueue::SRC_MUTE, id, mute ? 1.0f : 0.0f); });
}

int ssr_nu_set_master_volume(ssr_nu* r, float volume)
{
  return guarded([&] { if (!r) throw Error(SSR_ERR_INVALID, "ssr_b200: null renderer"); r->impl.push(ControlQueue::MASTER_VOLUME, 0, volume); });
}

int ssr_nu_sources(const ssr_nu* r) { return r ? r->impl.n_sources : 0; }
int ssr_nu_local_outputs(const ssr_nu* r) { return r ? r->impl.m_local : 0; }
long ssr_nu_taps_covered(const ssr_nu* r) { return r ? r->impl.taps_covered() : 0; }

int ssr_nu_process(ssr_nu* r, int n, const float* const* in, float* const* out)
{
  return guarded([&] {
    if (!r || (r->impl.n_sources > 0 && !in) || !out) throw Error(SSR_ERR_INVALID, "ssr_b200: nu_process: null argument");
    r->impl.process(n, in, out);
  });
}

int ssr_nu_process_device(ssr_nu* r, const float* d_in, float* d_out)
{
  return guarded([&] {
    if (!r || !d_out) throw Error(SSR_ERR_INVALID, "ssr_b200: nu_process_device: null argument");
    r->impl.step(d_in, d_out);
  });
}

int ssr_nu_synchronize(ssr_nu* r)
{
  return guarded([&] {
    if (!r) throw Error(SSR_ERR_INVALID, "ssr_b200: null renderer");
    SSR_CUDA(cudaSetDevice(r->impl.cfg.device));
    SSR_CUDA(cudaStreamSynchronize(r->impl.stream));
  });
}

void* ssr_nu_stream(ssr_nu* r) { return r ? static_cast<void*>(r->impl.stream) : nullptr; }

// kernels launched so far by all levels' engines and the mixer
int ssr_nu_launches(ssr_nu* r, uint64_t* launches)
{
  return guarded([&] {
    if (!r || !launches) throw Error(SSR_ERR_INVALID, "ssr_b200: null argument");
    uint64_t total = r->impl.mix_launches;
    for (auto& L : r->impl.levels) total += L->eng->stats.launches_total;
    *launches = total;
  });
}

// ALGORITHMIC bytes of the MAC launches of block `block_index` (SURVEY 8d formula per launch, summed):
// level 0 every block, level l's sub-renderer `block mod K` from the second period on.
int ssr_nu_block_mac_bytes(ssr_nu* r, uint64_t block_index, uint64_t* bytes)
{
  return guarded([&] {
    if (!r || !bytes) throw Error(SSR_ERR_INVALID, "ssr_b200: null argument");
    NuRenderer& R = r->impl;
    uint64_t total = 0;
    for (size_t l = 0; l < R.levels.size(); ++l)
    {
      auto& L = *R.levels[l];
      const uint64_t BL = uint64_t(L.K) * R.B;
      const int phase = int(block_index % uint64_t(L.K));
      if (l > 0 && (block_index / uint64_t(L.K) < 1 || phase >= int(L.subs.size()))) continue;
      Renderer& S = *L.subs[l == 0 ? 0 : size_t(phase)];
      uint64_t parts = 0;
      for (auto& s : S.sources) parts += uint64_t(s->irs->valid.empty() ? 0 : s->irs->valid[0]);
      total += 8 * BL * (parts * uint64_t(S.m_local) + parts + uint64_t(S.m_local));
    }
    *bytes = total;
  });
}

}  // extern "C"
