// ssr/renderers.h -- drop-in C++ adaptors for the block step of the reference's
// convolution renderers, on top of the B200 engine's C ABI (Level 2 of
// include/ssr_b200.h):
//
//   ssr::GenericRenderer   (src/genericrenderer.h)
//   ssr::BinauralRenderer  (src/binauralrenderer.h)
//   ssr::BrsRenderer       (src/brsrenderer.h)
//
// They keep the reference's names and per-block protocol as seen through
// apf::pointer_policy<float*> (apf/apf/pointer_policy.h:58-122) and
// RendererBase (src/rendererbase.h:84-103, 247-314, 419-423):
//
//   apf::parameter_map p; p.set("block_size", 1024); p.set("sample_rate", 48000);
//   ssr::BinauralRenderer r(p);          // + "hrir_file", "hrir_size"
//   r.load_reproduction_setup();
//   int id = r.add_source(src_params);   // "properties_file" for Generic / BRS
//   r.activate();
//   r.state.reference_orientation = Orientation(az);
//   r.get_source(id)->gain = 0.5f;
//   r.audio_callback(1024, in, out);
//
// Whole blocks are batched into ONE engine step (forward FFT -> MAC -> inverse
// FFT + crossfade + source sum); there is no per-pair convolve() call and no
// worker-thread fan-out ("threads" is accepted and ignored).
//
// Several GPUs in one process (GenericRenderer only): p.set("devices", "0,1,2,3")
// shards the outputs over one engine per listed device -- the counterpart of the
// reference spreading its Output items over worker threads
// (apf/apf/mimoprocessor.h:452-483).  Every control assignment goes to all shards,
// audio_callback() queues the block on the peer shards and runs it on the first
// one, whose inverse kernel's peers store their output blocks straight into its
// gather buffer over NVLink (include/ssr_b200.h, "Multi-GPU output gather").
//
// Out of scope and therefore different: GenericRenderer takes its loudspeaker
// count from the parameter "outputs" (the reference parses the reproduction-setup
// XML, src/loudspeakerrenderer.h:95-150); IR files are read by the small WAV
// reader below instead of libsndfile (PCM 8/16/24/32 and IEEE float 32/64).
#ifndef SSR_B200_SSR_RENDERERS_H
#define SSR_B200_SSR_RENDERERS_H

#include <cmath>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "ssr_b200.h"

#ifndef APF_PARAMETER_MAP_H
namespace apf
{
/// Minimal stand-in for apf::parameter_map (apf/apf/parameter_map.h): a
/// string -> string map with typed set/get.  If the reference header was
/// included first, that one is used instead.
class parameter_map
{
  public:
    template<typename T>
    const std::string& set(const std::string& k, const T& v)
    {
      std::ostringstream ss; ss << std::boolalpha << v;
      return _map[k] = ss.str();
    }
    bool has_key(const std::string& k) const { return _map.count(k) != 0; }
    const std::string& operator[](const std::string& k) const
    {
      auto it = _map.find(k);
      if (it == _map.end()) throw std::out_of_range("Key \"" + k + "\" doesn't exist in map!");
      return it->second;
    }
    template<typename T> T get(const std::string& k) const
    {
      std::istringstream ss((*this)[k]);
      T v = T();
      ss >> std::boolalpha >> v;
      if (ss.fail()) throw std::invalid_argument("S2RV(): Couldn't convert \"" + (*this)[k] + "\"!");
      return v;
    }
    template<typename T> T get(const std::string& k, const T& def) const
    {
      if (!has_key(k)) return def;
      std::istringstream ss((*this)[k]);
      T v = def;
      ss >> std::boolalpha >> v;
      return ss.fail() ? def : v;
    }
    std::string get(const std::string& k, const char* def) const
    { return has_key(k) ? (*this)[k] : std::string(def); }
  private:
    std::map<std::string, std::string> _map;
};
}  // namespace apf
#endif

#ifndef SSR_POSITION_H
struct Position { explicit Position(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} float x, y; };
#endif
#ifndef SSR_ORIENTATION_H
struct Orientation { explicit Orientation(float a = 0) : azimuth(a) {} float azimuth; };
#endif

namespace ssr
{

namespace b200
{
inline void check(int rc)
{
  if (rc == SSR_OK) return;
  const std::string msg = ssr_last_error();
  if (rc == SSR_ERR_CUDA || rc == SSR_ERR_NOMEM) throw std::runtime_error(msg);
  throw std::logic_error(msg);  // the reference throws std::logic_error for bad IR sets
}

/// Interleaved float samples of a WAV file (frames x channels).
struct SoundFile
{
  int channels = 0, samplerate = 0;
  long frames = 0;
  int format = 3, bits = 32;  // WAV format tag (1 PCM, 3 IEEE float) and sample width
  std::vector<float> data;
};

inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (uint32_t(p[3]) << 24); }
inline uint16_t rd16(const unsigned char* p) { return uint16_t(p[0] | (p[1] << 8)); }

/// Streaming WAV reader: the header is parsed on open(), frames are converted to float
/// as they are read (what libsndfile's SndfileHandle::readf does for the reference,
/// apf/apf/mimoprocessor_file_io.h:80-110) -- a long multichannel file is never held
/// in memory as a whole.
class WavReader
{
  public:
    WavReader() {}
    WavReader(const WavReader&) = delete;
    WavReader& operator=(const WavReader&) = delete;
    ~WavReader() { close(); }

    int channels = 0, samplerate = 0;
    long frames = 0;
    int format = 3, bits = 32;  // WAV format tag (1 PCM, 3 IEEE float) and sample width

    /// false: not a readable WAV file
    bool open(const std::string& name)
    {
      close();
      _f = std::fopen(name.c_str(), "rb");
      if (!_f) return false;
      unsigned char hdr[12];
      if (std::fread(hdr, 1, 12, _f) != 12 || std::memcmp(hdr, "RIFF", 4) || std::memcmp(hdr + 8, "WAVE", 4))
        return fail();
      bool have_fmt = false;
      for (;;)
      {
        unsigned char ch[8];
        if (std::fread(ch, 1, 8, _f) != 8) return fail();
        const uint32_t size = rd32(ch + 4);
        if (!std::memcmp(ch, "fmt ", 4))
        {
          std::vector<unsigned char> fmt(size);
          if (size < 16 || std::fread(fmt.data(), 1, size, _f) != size) return fail();
          format = rd16(fmt.data());
          channels = rd16(fmt.data() + 2);
          samplerate = int(rd32(fmt.data() + 4));
          bits = rd16(fmt.data() + 14);
          if (format == 0xFFFE && size >= 26) format = rd16(fmt.data() + 24);  // WAVE_FORMAT_EXTENSIBLE
          have_fmt = true;
          if (size & 1) std::fseek(_f, 1, SEEK_CUR);
        }
        else if (!std::memcmp(ch, "data", 4))
        {
          if (!have_fmt || channels <= 0 || bits % 8 != 0) return fail();
          const bool known = (format == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32))
              || (format == 3 && (bits == 32 || bits == 64));
          if (!known) return fail();
          // a truncated file delivers what is there (libsndfile does the same)
          const long here = std::ftell(_f);
          std::fseek(_f, 0, SEEK_END);
          const long avail = std::ftell(_f) - here;
          std::fseek(_f, here, SEEK_SET);
          const size_t bytes = std::min<size_t>(size, size_t(std::max<long>(avail, 0)));
          frames = long(bytes / (size_t(bits) / 8) / size_t(channels));
          _left = frames;
          return true;
        }
        else std::fseek(_f, long(size + (size & 1)), SEEK_CUR);
      }
    }

    /// reads up to `n` frames, interleaved (frames x channels) floats; returns the frames read
    long read(float* out, long n)
    {
      n = std::min(n, _left);
      if (n <= 0 || !_f) return 0;
      const size_t bps = size_t(bits) / 8;
      const size_t count = size_t(n) * channels;
      _raw.resize(count * bps);
      const size_t got = std::fread(_raw.data(), bps, count, _f) / size_t(channels);
      for (size_t i = 0; i < got * channels; ++i)
      {
        const unsigned char* p = _raw.data() + i * bps;
        float v = 0.0f;
        if (format == 1)  // PCM, scaled like libsndfile's float conversion
        {
          if (bits == 8) v = (float(p[0]) - 128.0f) / 128.0f;
          else if (bits == 16) v = float(int16_t(rd16(p))) / 32768.0f;
          else if (bits == 24) v = float(int32_t((p[0] << 8) | (p[1] << 16) | (uint32_t(p[2]) << 24)) >> 8) / 8388608.0f;
          else v = float(double(int32_t(rd32(p))) / 2147483648.0);
        }
        else
        {
          if (bits == 32) { uint32_t u = rd32(p); std::memcpy(&v, &u, 4); }
          else { double d; std::memcpy(&d, p, 8); v = float(d); }
        }
        out[i] = v;
      }
      _left -= long(got);
      return long(got);
    }

    void close() { if (_f) std::fclose(_f); _f = nullptr; _left = 0; }

  private:
    bool fail() { close(); return false; }
    std::FILE* _f = nullptr;
    long _left = 0;
    std::vector<unsigned char> _raw;
};

/// Replaces apf::load_sndfile (apf/apf/sndfiletools.h:43-90), same checks and
/// std::logic_error messages.  (Impulse-response sets are loaded as a whole, as in the
/// reference.)
inline SoundFile load_sndfile(const std::string& name, size_t sample_rate, size_t channels)
{
  if (name == "") throw std::logic_error("apf::load_sndfile(): Empty file name!");
  WavReader rd;
  if (!rd.open(name) || !rd.channels)
    throw std::logic_error("apf::load_sndfile(): \"" + name + "\" couldn't be loaded!");
  SoundFile sf;
  sf.channels = rd.channels; sf.samplerate = rd.samplerate; sf.format = rd.format; sf.bits = rd.bits;
  sf.data.resize(size_t(rd.frames) * rd.channels);
  sf.frames = rd.read(sf.data.data(), rd.frames);
  sf.data.resize(size_t(sf.frames) * rd.channels);
  if (sample_rate && sample_rate != size_t(sf.samplerate))
    throw std::logic_error("apf::load_sndfile(): \"" + name + "\" has sample rate "
        + std::to_string(sf.samplerate) + " instead of " + std::to_string(sample_rate) + "!");
  if (channels && channels != size_t(sf.channels))
    throw std::logic_error("apf::load_sndfile(): \"" + name + "\" has "
        + std::to_string(sf.channels) + " channels instead of " + std::to_string(channels) + "!");
  return sf;
}

/// Streaming WAV writer (PCM 16/24/32 or IEEE float 32): header first, frames as they
/// come, sizes patched on close().
class WavWriter
{
  public:
    WavWriter() {}
    WavWriter(const WavWriter&) = delete;
    WavWriter& operator=(const WavWriter&) = delete;
    ~WavWriter() { close(); }

    void open(const std::string& name, int channels, int samplerate, int format = 3, int bits = 32)
    {
      close();
      if (!((format == 1 && (bits == 16 || bits == 24 || bits == 32)) || (format == 3 && bits == 32)))
      { format = 3; bits = 32; }
      _f = std::fopen(name.c_str(), "wb");
      if (!_f) throw std::runtime_error("write_wav(): cannot open \"" + name + "\"");
      _channels = channels; _format = format; _bits = bits; _bytes = 0;
      const uint32_t bps = uint32_t(bits) / 8;
      std::fwrite("RIFF", 1, 4, _f); put32(36); std::fwrite("WAVE", 1, 4, _f);
      std::fwrite("fmt ", 1, 4, _f); put32(16); put16(uint16_t(format)); put16(uint16_t(channels));
      put32(uint32_t(samplerate)); put32(uint32_t(samplerate) * channels * bps);
      put16(uint16_t(channels * bps)); put16(uint16_t(bits));
      std::fwrite("data", 1, 4, _f); put32(0);
    }

    void write(const float* data, long frames)
    {
      if (!_f || frames <= 0) return;
      const size_t n = size_t(frames) * _channels;
      const uint32_t bps = uint32_t(_bits) / 8;
      if (_format == 3) std::fwrite(data, 4, n, _f);
      else
      {
        const double full = _bits == 16 ? 32768.0 : (_bits == 24 ? 8388608.0 : 2147483648.0);
        _raw.resize(n * bps);
        for (size_t i = 0; i < n; ++i)
        {
          double v = std::floor(double(data[i]) * full + 0.5);
          if (v > full - 1) v = full - 1;
          if (v < -full) v = -full;
          const int32_t q = int32_t(v);
          std::memcpy(_raw.data() + i * bps, &q, bps);  // little endian host
        }
        std::fwrite(_raw.data(), 1, _raw.size(), _f);
      }
      _bytes += uint64_t(n) * bps;
    }

    void close()
    {
      if (!_f) return;
      const uint32_t bytes = uint32_t(_bytes);
      std::fseek(_f, 4, SEEK_SET); put32(36 + bytes);
      std::fseek(_f, 40, SEEK_SET); put32(bytes);
      std::fclose(_f);
      _f = nullptr;
    }

  private:
    void put32(uint32_t v) { std::fwrite(&v, 4, 1, _f); }
    void put16(uint16_t v) { std::fwrite(&v, 2, 1, _f); }
    std::FILE* _f = nullptr;
    int _channels = 0, _format = 3, _bits = 32;
    uint64_t _bytes = 0;
    std::vector<unsigned char> _raw;
};

/// Writes interleaved float samples as WAV in one go.
inline void write_wav(const std::string& name, const float* data, long frames, int channels
    , int samplerate, int format = 3, int bits = 32)
{
  WavWriter w;
  w.open(name, channels, samplerate, format, bits);
  w.write(data, frames);
  w.close();
}

/// Assignable control value forwarding to a C setter (stands in for
/// apf::SharedData<T>, apf/apf/shareddata.h:35-83: assignment takes effect at the
/// start of the next audio_callback).
template<typename T, typename F>
class Shared
{
  public:
    Shared(T init, F f) : _v(init), _f(f) {}
    Shared& operator=(const T& v) { _v = v; _f(v); return *this; }
    operator const T&() const { return _v; }
    const T& get() const { return _v; }
  private:
    T _v; F _f;
};

enum model_t { point = 0, plane = 1 };

class RendererCore
{
  public:
    class Source
    {
      public:
        Source(RendererCore* c, int id_)
          : id(id_)
          , position(Position(), [c, id_](const Position& p) { c->each([&](ssr_renderer* r) { check(ssr_source_set_position(r, id_, p.x, p.y)); }); })
          , orientation(Orientation(), [](const Orientation&) {})
          , gain(1.0f, [c, id_](float g) { c->each([&](ssr_renderer* r) { check(ssr_source_set_gain(r, id_, g)); }); })
          , mute(false, [c, id_](bool m) { c->each([&](ssr_renderer* r) { check(ssr_source_set_mute(r, id_, m)); }); })
          , model(point, [c, id_](model_t m) { c->each([&](ssr_renderer* r) { check(ssr_source_set_model(r, id_, m == plane)); }); })
        {}
        const int id;
        Shared<Position, std::function<void(const Position&)>> position;
        Shared<Orientation, std::function<void(const Orientation&)>> orientation;
        Shared<float, std::function<void(float)>> gain;
        Shared<bool, std::function<void(bool)>> mute;
        Shared<model_t, std::function<void(model_t)>> model;
    };

    struct State
    {
      explicit State(RendererCore* c)
        : reference_position(Position(), [c](const Position& p) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_reference_position(r, p.x, p.y)); }); })
        , reference_orientation(Orientation(), [c](const Orientation& o) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_reference_orientation(r, o.azimuth)); }); })
        , reference_offset_position(Position(), [c](const Position& p) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_reference_offset_position(r, p.x, p.y)); }); })
        , reference_offset_orientation(Orientation(), [c](const Orientation& o) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_reference_offset_orientation(r, o.azimuth)); }); })
        , master_volume(1.0f, [c](float v) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_master_volume(r, v)); }); })
        , processing(true, [c](bool v) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_processing(r, v)); }); })
        , amplitude_reference_distance(3.0f, [c](float v) { c->each([&](ssr_renderer* r) { check(ssr_renderer_set_amplitude_reference_distance(r, v)); }); })
      {}
      Shared<Position, std::function<void(const Position&)>> reference_position;
      Shared<Orientation, std::function<void(const Orientation&)>> reference_orientation;
      Shared<Position, std::function<void(const Position&)>> reference_offset_position;
      Shared<Orientation, std::function<void(const Orientation&)>> reference_offset_orientation;
      Shared<float, std::function<void(float)>> master_volume;
      Shared<bool, std::function<void(bool)>> processing;
      Shared<float, std::function<void(float)>> amplitude_reference_distance;
    };

    int block_size() const { return _block_size; }
    int sample_rate() const { return _sample_rate; }
    int in_channels() const { return ssr_renderer_sources(_r); }
    /// outputs audio_callback() delivers (all shards' outputs when sharded)
    int out_channels() const { return _gathers.empty() ? ssr_renderer_local_outputs(_r) : ssr_gather_total_outputs(_gathers[0]); }
    int shards() const { return 1 + int(_peer_r.size()); }
    /// can audio_callback_batch() share passes over the filters between blocks?
    bool batchable() const { return _peer_r.empty() && _kind == SSR_RENDERER_GENERIC && _block_size >= 1024; }

    /// f(renderer handle) on every shard (one shard unless "devices" lists several)
    template<typename F> void each(F f) { f(_r); for (auto r : _peer_r) f(r); }

    bool activate() { return true; }
    bool deactivate() { return true; }

    /// = pointer_policy::audio_callback (apf/apf/pointer_policy.h:112-122)
    void audio_callback(int n, float* const* in, float* const* out)
    {
      // sharded: the peers queue the block (their inverse kernels deliver into the
      // first shard's gather buffer), the first shard runs it and returns all outputs
      for (auto r : _peer_r) check(ssr_renderer_submit(r, n, in));
      check(ssr_renderer_process(_r, n, in, out));
      for (auto r : _peer_r) check(ssr_renderer_wait(r, nullptr));
    }

    /// Offline: `nblocks` consecutive blocks in one call (in[t * in_channels() + i],
    /// out[t * out_channels() + m]); runs of constant-control blocks share one pass
    /// over the filter spectra (ssr_renderer_process_batch).  Unsharded renderers only.
    void audio_callback_batch(int n, int nblocks, const float* const* in, float* const* out)
    {
      if (!_peer_r.empty()) throw std::logic_error("ssr_b200: audio_callback_batch on a sharded renderer");
      check(ssr_renderer_process_batch(_r, n, nblocks, in, out));
    }

    /// Pipelined form of audio_callback (offline rendering): submit() queues a block
    /// and returns, wait() delivers the oldest queued block; at most two in flight.
    void submit(int n, const float* const* in)
    {
      for (auto r : _peer_r) check(ssr_renderer_submit(r, n, in));
      check(ssr_renderer_submit(_r, n, in));
    }
    void wait(float* const* out)
    {
      check(ssr_renderer_wait(_r, out));
      for (auto r : _peer_r) check(ssr_renderer_wait(r, nullptr));
    }

    Source* get_source(int id)
    {
      auto it = _sources.find(id);
      return it == _sources.end() ? nullptr : it->second.get();
    }

    void rem_source(Source* s)
    {
      if (!s) return;
      const int id = s->id;
      each([&](ssr_renderer* r) { check(ssr_renderer_rem_source(r, id)); });
      _sources.erase(id);
    }

    void rem_all_sources() { while (!_sources.empty()) rem_source(_sources.begin()->second.get()); }

    const apf::parameter_map params;
    State state;

    ssr_renderer* handle() const { return _r; }
    ssr_engine* engine() const { return _e; }

    RendererCore(const RendererCore&) = delete;
    RendererCore& operator=(const RendererCore&) = delete;

  protected:
    RendererCore(const apf::parameter_map& p, ssr_renderer_kind kind)
      : params(p), state(this)
      , _block_size(p.get<int>("block_size")), _sample_rate(p.get<int>("sample_rate"))
      , _kind(kind), _e(nullptr), _r(nullptr)
    {
      // "devices": comma-separated CUDA ordinals, one output shard per entry
      std::vector<int> devices;
      {
        std::stringstream ss(p.get("devices", std::string()));
        std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) devices.push_back(std::atoi(tok.c_str()));
      }
      if (devices.empty()) devices.push_back(p.get("device", 0));
      const int world = int(devices.size());
      if (world > 1 && kind != SSR_RENDERER_GENERIC)
        throw std::logic_error("ssr_b200: only the GenericRenderer shards over several devices");
      const int outputs = kind == SSR_RENDERER_GENERIC ? p.get("outputs", 0) : 2;
      if (world > 1 && outputs < world) throw std::logic_error("ssr_b200: fewer outputs than devices");
      std::vector<int> counts(size_t(world), outputs / world);
      for (int g = 0; g < outputs % world; ++g) counts[size_t(g)]++;
      try
      {
        int first = 0;
        for (int g = 0; g < world; ++g)
        {
          ssr_engine_config ec = ssr_engine_config();
          ec.device = devices[size_t(g)];
          ec.block_size = _block_size;
          ec.sample_rate = _sample_rate;
          ssr_engine* e = nullptr;
          check(ssr_engine_create(&ec, &e));
          (g == 0 ? _e : *_peer_e.insert(_peer_e.end(), nullptr)) = e;
          ssr_renderer_config rc = ssr_renderer_config();
          rc.kind = kind;
          rc.outputs = outputs;
          rc.output_first = world > 1 ? first : p.get("output_first", 0);
          rc.output_count = world > 1 ? counts[size_t(g)] : p.get("output_count", 0);
          rc.master_volume_correction_db = float(p.get("master_volume_correction", 0.0));
          ssr_renderer* r = nullptr;
          check(ssr_renderer_create(e, &rc, &r));
          (g == 0 ? _r : *_peer_r.insert(_peer_r.end(), nullptr)) = r;
          first += counts[size_t(g)];
        }
        if (world > 1)
        {
          for (int g = 0; g < world; ++g)
          {
            ssr_gather* ga = nullptr;
            check(ssr_gather_create(g == 0 ? _e : _peer_e[size_t(g - 1)], world, g, 0, counts.data(), &ga));
            _gathers.push_back(ga);
            if (g > 0) check(ssr_gather_connect_local(ga, _gathers[0]));
          }
          for (int g = 0; g < world; ++g)
            check(ssr_renderer_bind_gather(g == 0 ? _r : _peer_r[size_t(g - 1)], _gathers[size_t(g)]));
        }
      }
      catch (...) { destroy(); throw; }
    }

    ~RendererCore() { destroy(); }

    int register_source(int id)
    {
      _sources[id].reset(new Source(this, id));
      return id;
    }

    /// ssr_renderer_add_source on every shard (each picks its own output columns);
    /// the shards hand out the same id
    int add_source_all(const float* ir, long frames, int channels)
    {
      int id = 0;
      check(ssr_renderer_add_source(_r, ir, frames, channels, &id));
      for (auto r : _peer_r)
      {
        int pid = 0;
        check(ssr_renderer_add_source(r, ir, frames, channels, &pid));
        if (pid != id) throw std::logic_error("ssr_b200: shards disagree on a source id");
      }
      return id;
    }

    const int _block_size, _sample_rate;
    const ssr_renderer_kind _kind;
    ssr_engine* _e;
    ssr_renderer* _r;                       // the first (host-facing) shard
    std::vector<ssr_engine*> _peer_e;       // further shards ("devices")
    std::vector<ssr_renderer*> _peer_r;
    std::vector<ssr_gather*> _gathers;
    std::map<int, std::unique_ptr<Source>> _sources;

  private:
    void destroy()
    {
      _sources.clear();
      // renderers first (they only point at their gathers), the peers' gathers before
      // the root's buffer they map, engines last
      for (auto r : _peer_r) if (r) ssr_renderer_destroy(r);
      if (_r) ssr_renderer_destroy(_r);
      for (size_t g = _gathers.size(); g-- > 0;) if (_gathers[g]) ssr_gather_destroy(_gathers[g]);
      for (auto e : _peer_e) if (e) ssr_engine_destroy(e);
      if (_e) ssr_engine_destroy(_e);
      _peer_r.clear(); _peer_e.clear(); _gathers.clear(); _r = nullptr; _e = nullptr;
    }
};
}  // namespace b200

class GenericRenderer : public b200::RendererCore
{
  public:
    static const char* name() { return "GenericRenderer"; }
    explicit GenericRenderer(const apf::parameter_map& p) : RendererCore(p, SSR_RENDERER_GENERIC) {}
    void load_reproduction_setup() {}  // loudspeaker count comes from "outputs"

    /// = add_source with GenericRenderer::Source (src/genericrenderer.h:88-120):
    /// "properties_file" must have exactly as many channels as there are outputs.
    int add_source(const apf::parameter_map& p = apf::parameter_map())
    {
      auto ir = b200::load_sndfile(p.get("properties_file", ""), size_t(sample_rate())
          , size_t(params.get("outputs", 0)));
      return register_source(add_source_all(ir.data.data(), long(ir.frames), int(ir.channels)));
    }
};

class BrsRenderer : public b200::RendererCore
{
  public:
    static const char* name() { return "BrsRenderer"; }
    explicit BrsRenderer(const apf::parameter_map& p) : RendererCore(p, SSR_RENDERER_BRS) {}
    void load_reproduction_setup() {}  // two outputs (src/brsrenderer.h:278-301)

    /// = add_source with BrsRenderer::Source (src/brsrenderer.h:90-139)
    int add_source(const apf::parameter_map& p = apf::parameter_map())
    {
      auto ir = b200::load_sndfile(p.get("properties_file", ""), size_t(sample_rate()), 0);
      int id = 0;
      b200::check(ssr_renderer_add_source(_r, ir.data.data(), ir.frames, ir.channels, &id));
      return register_source(id);
    }
};

class BinauralRenderer : public b200::RendererCore
{
  public:
    static const char* name() { return "BinauralRenderer"; }
    explicit BinauralRenderer(const apf::parameter_map& p) : RendererCore(p, SSR_RENDERER_BINAURAL) {}

    /// = BinauralRenderer::load_reproduction_setup -> _load_hrtfs
    /// (src/binauralrenderer.h:117-169, 211-233)
    void load_reproduction_setup()
    {
      auto hrirs = b200::load_sndfile(params["hrir_file"], size_t(sample_rate()), 0);
      b200::check(ssr_renderer_load_hrirs(_r, hrirs.data.data(), hrirs.frames, hrirs.channels
            , long(params.get("hrir_size", 0))));
    }

    int add_source(const apf::parameter_map& = apf::parameter_map())
    {
      int id = 0;
      b200::check(ssr_renderer_add_source(_r, nullptr, 0, 0, &id));
      return register_source(id);
    }
};

/// The pre-filter stage of the WFS renderer (src/wfsrenderer.h:60-123): the
/// renderer loads "prefilter_file" (single channel) and every Input owns a
/// StaticConvolver on it whose output feeds the (out-of-scope) delay line.
/// Here all inputs are convolved in one batched step per block.
class WfsPrefilterBank : public b200::RendererCore
{
  public:
    static const char* name() { return "WfsPrefilterBank"; }
    explicit WfsPrefilterBank(const apf::parameter_map& p)
      : RendererCore(p, SSR_RENDERER_PREFILTER_BANK)
    {
      // load_sndfile(prefilter_file, sample_rate, 1) (src/wfsrenderer.h:75-76)
      auto pf = b200::load_sndfile(params.get("prefilter_file", ""), size_t(sample_rate()), 1);
      b200::check(ssr_renderer_load_prefilter(_r, pf.data.data(), pf.frames));
    }

    /// adds one input; its pre-filtered signal is output number in_channels() - 1
    int add_input()
    {
      int id = 0;
      b200::check(ssr_renderer_add_source(_r, nullptr, 0, 0, &id));
      return id;
    }
};

}  // namespace ssr

#endif
