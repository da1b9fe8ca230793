// ssr_b200_render -- offline rendering of a multichannel WAV file through one of
// the GPU renderers (the reference's mimoprocessor_file_io use case,
// apf/apf/mimoprocessor_file_io.h:40-117).
//
//   ssr_b200_render generic  <block> <in.wav> <out.wav> <ir_src1.wav> [<ir_src2.wav> ...]
//   ssr_b200_render brs      <block> <in.wav> <out.wav> <brir_src1.wav> [...]  [--azimuth deg]
//   ssr_b200_render binaural <block> <in.wav> <out.wav> <hrirs.wav> [--azimuth deg]
//
// One IR file per input channel (source); Generic IR files have one channel per
// loudspeaker.  Binaural sources are placed on a 2 m circle around the listener.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "apf/mimoprocessor_file_io.h"

template<typename R>
static int run(R& r, const std::string& in, const std::string& out)
{
  return apf::mimoprocessor_file_io(r, in, out);
}

int main(int argc, char* argv[])
{
  if (argc < 6)
  {
    std::fprintf(stderr, "usage: %s generic|brs|binaural <block> <in.wav> <out.wav> <ir.wav>... [--azimuth deg] [--devices 0,1,...]\n", argv[0]);
    return 2;
  }
  const std::string kind = argv[1];
  const int block = std::atoi(argv[2]);
  const std::string in = argv[3], out = argv[4];
  std::vector<std::string> files;
  float azimuth = 0.0f;
  bool have_az = false;
  std::string devices;   // generic only: comma-separated CUDA ordinals, one output shard each
  for (int i = 5; i < argc; ++i)
  {
    if (!std::strcmp(argv[i], "--azimuth") && i + 1 < argc) { azimuth = float(std::atof(argv[++i])); have_az = true; }
    else if (!std::strcmp(argv[i], "--devices") && i + 1 < argc) devices = argv[++i];
    else files.push_back(argv[i]);
  }
  try
  {
    auto input = ssr::b200::load_sndfile(in, 0, 0);
    apf::parameter_map p;
    p.set("block_size", block);
    p.set("sample_rate", input.samplerate);
    if (kind == "generic")
    {
      auto first = ssr::b200::load_sndfile(files.at(0), 0, 0);
      p.set("outputs", first.channels);
      if (!devices.empty()) p.set("devices", devices);
      ssr::GenericRenderer r(p);
      for (auto& f : files) { apf::parameter_map sp; sp.set("properties_file", f); r.add_source(sp); }
      return run(r, in, out);
    }
    if (kind == "brs")
    {
      ssr::BrsRenderer r(p);
      r.load_reproduction_setup();
      for (auto& f : files) { apf::parameter_map sp; sp.set("properties_file", f); r.add_source(sp); }
      if (have_az) r.state.reference_orientation = Orientation(azimuth);
      return run(r, in, out);
    }
    if (kind == "binaural")
    {
      p.set("hrir_file", files.at(0));
      ssr::BinauralRenderer r(p);
      r.load_reproduction_setup();
      for (int i = 0; i < input.channels; ++i)
      {
        const int id = r.add_source();
        const double a = 2.0 * M_PI * i / input.channels;
        r.get_source(id)->position = Position(float(2 * std::cos(a)), float(2 * std::sin(a)));
      }
      if (have_az) r.state.reference_orientation = Orientation(azimuth);
      return run(r, in, out);
    }
    std::fprintf(stderr, "unknown renderer kind %s\n", kind.c_str());
    return 2;
  }
  catch (const std::exception& e)
  {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
