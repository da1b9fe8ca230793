// hostgather.cu -- host-direct output delivery for a sharded renderer (see engine.hpp,
// struct HostGather): every rank's inverse kernel stores its output blocks straight
// into ONE host memory segment shared by all ranks (POSIX shared memory, page-locked
// and mapped into every rank's device address space), over that rank's own PCIe
// link.  Nothing passes through the host-facing rank's GPU and no rank ever spins
// on a device: all waiting is done by host threads on flags in the segment.
#include "engine.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>

namespace ssr_b200
{

using namespace ssrk;

static inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

HostGather::HostGather(Engine* e_, int world_, int rank_, int root_, const int* outputs_per_rank, const char* shm_name)
  : e(e_), world(world_), rank(rank_), root(root_), name(shm_name ? shm_name : "")
{
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || root < 0 || root >= world || !outputs_per_rank)
    throw Error(SSR_ERR_INVALID, "ssr_b200: hostgather: bad world / rank / root");
  if (name.empty() || name[0] != '/' || name.find('/', 1) != std::string::npos)
    throw Error(SSR_ERR_INVALID, "ssr_b200: hostgather: the segment name must look like \"/name\" (shm_open)");
  for (int g = 0; g < world; ++g)
  {
    if (outputs_per_rank[g] < 0) throw Error(SSR_ERR_INVALID, "ssr_b200: hostgather: negative output count");
    firsts.push_back(M_total);
    counts.push_back(outputs_per_rank[g]);
    M_total += outputs_per_rank[g];
  }
  if (M_total == 0) throw Error(SSR_ERR_INVALID, "ssr_b200: hostgather: no outputs");
  if (const char* v = std::getenv("SSR_B200_GATHER_TIMEOUT_MS")) timeout_ms = std::max(1.0, std::atof(v));
  e->use_device();
  bytes = kControlBytes + 2 * slot_floats() * sizeof(float);

  int fd = -1;
  if (is_root())
  {
    fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) throw Error(SSR_ERR_STATE, "ssr_b200: hostgather: shm_open(" + name + ", create) failed: " + std::strerror(errno));
    created = true;
    if (ftruncate(fd, off_t(bytes)) != 0)
    {
      const std::string why = std::strerror(errno);
      close(fd); shm_unlink(name.c_str()); created = false;
      throw Error(SSR_ERR_NOMEM, "ssr_b200: hostgather: ftruncate failed: " + why);
    }
  }
  else
  {
    // the root creates the segment; wait until it exists at its full size
    const auto t0 = std::chrono::steady_clock::now();
    while (true)
    {
      fd = shm_open(name.c_str(), O_RDWR, 0600);
      if (fd >= 0)
      {
        struct stat st;
        if (fstat(fd, &st) == 0 && size_t(st.st_size) >= bytes) break;
        close(fd); fd = -1;
      }
      if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms)
        throw Error(SSR_ERR_STATE, "ssr_b200: hostgather: segment " + name + " did not appear (is the root rank up?)");
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED)
  {
    if (created) { shm_unlink(name.c_str()); created = false; }
    throw Error(SSR_ERR_NOMEM, std::string("ssr_b200: hostgather: mmap failed: ") + std::strerror(errno));
  }
  base = static_cast<unsigned char*>(p);
  if (is_root()) std::memset(base, 0, bytes);   // also faults every page in before it is pinned
  cudaError_t err = cudaHostRegister(base, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (err != cudaSuccess)
  {
    munmap(base, bytes); base = nullptr;
    if (created) { shm_unlink(name.c_str()); created = false; }
    throw Error(SSR_ERR_CUDA, std::string("cudaHostRegister (hostgather segment): ") + cudaGetErrorString(err));
  }
  registered = true;
  void* dp = nullptr;
  SSR_CUDA(cudaHostGetDevicePointer(&dp, base, 0));
  d_base = static_cast<unsigned char*>(dp);
  d_counter.resize(1);
  SSR_CUDA(cudaMemset(d_counter.ptr, 0, sizeof(unsigned)));
}

HostGather::~HostGather()
{
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (registered) cudaHostUnregister(base);
  if (base) munmap(base, bytes);
  if (created) shm_unlink(name.c_str());
}

// kernel arguments of the block about to be rendered; then ++epoch
GatherArgs HostGather::next_block(int ctas)
{
  GatherArgs ga{};
  const uint64_t t = epoch++;
  ga.out_offset = (long long)((t & 1) * slot_floats());
  ga.cta_counter = d_counter.ptr;
  ga.cta_total = unsigned(ctas);
  ga.host_flag = reinterpret_cast<unsigned long long*>(d_base) + size_t(1 + rank) * kGatherCounterStride;
  ga.flag_value = t + 1;
  return ga;
}

static bool spin_until_ge(volatile unsigned long long* p, unsigned long long target, double timeout_ms)
{
  if (*p >= target) return true;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 1; ; ++spins)
  {
    if (*p >= target) return true;
    cpu_relax();
    if ((spins & 0x3ffu) == 0
        && std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms)
      return false;
  }
}

// root, at the start of block t: the caller is done with every block up to t - 2
// (slot t & 1 is about to be overwritten)
void HostGather::publish_consumed(uint64_t t)
{
  if (t >= 1) __atomic_store_n(consumed_ptr(), (unsigned long long)(t - 1), __ATOMIC_RELEASE);
}

// peer, before the stores of block t are launched: slot t & 1 has been read out
void HostGather::wait_slot_free(uint64_t t)
{
  if (is_root() || t < 2) return;
  if (!spin_until_ge(consumed_ptr(), t - 1, timeout_ms))
  {
    failed = true;
    throw Error(SSR_ERR_STATE, "ssr_b200: hostgather: the root rank did not release the output slot in time");
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
}

// root: every rank's blocks of block t have landed in the segment
void HostGather::wait_all(uint64_t t)
{
  for (int g = 0; g < world; ++g)
  {
    if (counts[g] == 0) continue;
    if (!spin_until_ge(flag_ptr(g), t + 1, timeout_ms))
    {
      failed = true;
      throw Error(SSR_ERR_STATE, "ssr_b200: hostgather: rank " + std::to_string(g) + " did not deliver its output blocks in time");
    }
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
}

}  // namespace ssr_b200
