// CPU-only test (no GPU, no CUDA call) of the engine's pure host logic:
//  * FilterPtrTable = Output::set_filter / queues_empty / rotate_queues
//    (apf/apf/convolver.h:618-699), checked against a literal restatement of the
//    reference's shifting queues and against the sequence the reference's unit
//    test expects (apf/unit_tests/test_convolver.cpp:86-131);
//  * MacPlan::make_work: every term of every group is covered exactly once by
//    contiguous chunks, partial indices are unique, chunks are balanced;
//  * the rule the dynamic renderers' device-side term generation rests on
//    (mac.cuh, DynTables): with the renderers' protocol (old state, rotate
//    while the queues are not empty, set_filter), in block b partition p uses the
//    filter that was current in block b - p, the state before the switch is the
//    previous block's state, and queues_empty() <=> b >= b_set + P;
//  * ssrk::dyn_split: the CTA ranges of the three accumulator groups.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <set>
#include <thread>
#include <vector>

#include "../../ssr-pre-0.4.0_b200/csrc/engine.hpp"

using namespace ssr_b200;

static int failures = 0, checks = 0;
#define CHECK(cond) do { ++checks; if (!(cond)) { ++failures; \
  std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)

// literal restatement of the reference's data structure: _queues[i] has i+1
// slots and is shifted by one on every rotate (convolver.h:623-624, 642-699).
// The reference's "_empty_partition" is a valid non-null pointer; here it is a
// distinguished address.
struct LiteralQueues
{
  explicit LiteralQueues(int P_, const float2* empty_) : P(P_), empty(empty_), ptrs(P_, empty_)
  { for (int i = 0; i + 1 < P; ++i) queues.emplace_back(size_t(i + 1), nullptr); }
  void set_filter(const std::vector<const float2*>& f)
  {
    size_t k = 0;
    if (k < f.size()) ptrs[0] = f[k++];
    for (size_t i = 0; i < queues.size(); ++i) queues[i][i] = (k >= f.size()) ? empty : f[k++];
  }
  bool queues_empty() const
  {
    if (queues.empty()) return true;
    for (auto p : queues.back()) if (p) return false;
    return true;
  }
  void rotate_queues()
  {
    for (size_t i = 0; i < queues.size(); ++i)
    {
      auto& q = queues[i];
      if (q.front()) ptrs[i + 1] = q.front();
      for (size_t j = 0; j + 1 < q.size(); ++j) q[j] = q[j + 1];
      q.back() = nullptr;
    }
  }
  int P; const float2* empty;
  std::vector<const float2*> ptrs;
  std::vector<std::vector<const float2*>> queues;
};

static unsigned rng_state = 4711u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static void test_ptr_table()
{
  static float2 storage[64][64];  // fake partition addresses: filter f, partition p
  static float2 empty_marker;
  for (int P : {1, 2, 3, 5, 8, 47})
  {
    FilterPtrTable t(P);
    LiteralQueues lit(P, &empty_marker);
    CHECK(t.queues_empty());
    for (int step = 0; step < 400; ++step)
    {
      const unsigned op = rnd() % 10;
      if (op < 3)
      {
        const int f = int(rnd() % 64);
        const int len = 1 + int(rnd() % (P + 2));  // shorter, equal or longer than P
        std::vector<const float2*> parts;
        for (int p = 0; p < len && p < 64; ++p) parts.push_back(&storage[f][p]);
        t.set_filter_ptrs(parts);
        lit.set_filter(parts);
      }
      else if (op < 8)
      {
        // the renderers rotate only while the queues are not empty
        if (!lit.queues_empty()) { t.rotate_queues(); lit.rotate_queues(); }
      }
      else { t.rotate_queues(); lit.rotate_queues(); }  // rotating empty queues is harmless
      CHECK(t.queues_empty() == lit.queues_empty());
      bool same = true;
      for (int p = 0; p < P; ++p)
      {
        const float2* want = lit.ptrs[p] == &empty_marker ? nullptr : lit.ptrs[p];
        same = same && (t.current()[p] == want);
      }
      CHECK(same);
    }
  }

  // the sequence of test_convolver.cpp:86-131 (P = 2)
  FilterPtrTable t(2);
  std::vector<const float2*> f = { &storage[1][0], &storage[1][1] };
  t.set_filter_ptrs(f);
  CHECK(t.current()[0] == &storage[1][0] && t.current()[1] == nullptr);
  CHECK(!t.queues_empty());
  t.rotate_queues();
  CHECK(t.current()[1] == &storage[1][1]);
  CHECK(t.queues_empty());

  // StaticOutput::_set_filter: fewer partitions -> rest empty, more -> ignored
  FilterPtrTable s(3);
  s.set_static_ptrs({ &storage[2][0] });
  CHECK(s.current()[0] == &storage[2][0] && s.current()[1] == nullptr && s.current()[2] == nullptr);
  s.set_static_ptrs({ &storage[3][0], &storage[3][1], &storage[3][2], &storage[3][3] });
  CHECK(s.current()[2] == &storage[3][2]);
  const uint64_t v = s.version;
  s.rotate_queues();
  CHECK(s.version != v);
}

static void test_make_work()
{
  for (int trial = 0; trial < 300; ++trial)
  {
    MacPlan plan;
    plan.MT = 1 << (rnd() % 3);
    const int G = 1 + int(rnd() % 140);
    const bool uniform = rnd() % 2;
    const int base = 1 + int(rnd() % 7000);
    long total = 0;
    const float4* h[4] = { nullptr, nullptr, nullptr, nullptr };
    for (int g = 0; g < G; ++g)
    {
      plan.begin_group();
      const int n = uniform ? base : int(rnd() % (base + 1));  // empty groups allowed
      for (int i = 0; i < n && total < 200000; ++i, ++total) plan.add_term(g, i, 0, h);
      plan.end_group();
    }
    const int slots = 148 * (1 + int(rnd() % 2));
    const int fixed = (rnd() % 2) ? 2 : 8;
    const int waves = int(rnd() % 3);   // 0 = default (one wave)
    plan.make_work(slots, fixed, waves);
    CHECK(int(plan.work.host.size()) == plan.npartials);
    CHECK(int(plan.cta_first.host.size()) == plan.ncta + 1);
    CHECK(plan.ncta >= 1 && plan.ncta <= slots * std::max(1, waves));
    // every segment lies inside one group, carries a unique partial index, and the
    // segments of a group cover its terms exactly once with CONSECUTIVE partial indices
    std::vector<int4> by_partial(plan.work.host.size());
    std::set<int> partial_ids;
    bool ok = true;
    for (auto& item : plan.work.host)
    {
      ok = ok && item.z >= 0 && item.z < plan.npartials && item.y > item.x;
      if (!ok) break;
      partial_ids.insert(item.z);
      by_partial[size_t(item.z)] = item;
    }
    CHECK(ok);
    CHECK(int(partial_ids.size()) == plan.npartials);
    for (auto& g : plan.groups)
    {
      int pos = g.term_begin;
      ok = ok && (g.nsplit >= 1 || g.term_end == g.term_begin);
      for (int s = 0; ok && s < g.nsplit; ++s)
      {
        const int4 item = by_partial[size_t(g.partial_first + s)];
        ok = ok && item.x == pos && item.y <= g.term_end;
        pos = item.y;
      }
      ok = ok && pos == g.term_end;
    }
    CHECK(ok);
    // CTA shares: contiguous, balanced to within one term, longest segment first
    long lo = 1 << 30, hi = 0, covered = 0;
    for (int c = 0; c < plan.ncta; ++c)
    {
      const int a = plan.cta_first.host[size_t(c)], b = plan.cta_first.host[size_t(c) + 1];
      ok = ok && a <= b;
      long share = 0;
      int first_len = 0, max_len = 0;
      for (int i = a; i < b; ++i)
      {
        const int len = plan.work.host[size_t(i)].y - plan.work.host[size_t(i)].x;
        if (i == a) first_len = len;
        max_len = std::max(max_len, len);
        share += len;
      }
      ok = ok && first_len == max_len;
      lo = std::min(lo, share); hi = std::max(hi, share); covered += share;
    }
    CHECK(ok);
    CHECK(covered == total);
    if (total >= plan.ncta) CHECK(hi - lo <= 1);
    // no pathological over-splitting: a CTA adds at most one segment per group it touches
    CHECK(plan.npartials <= plan.ncta + G);
  }
}

// One Output driven like BrsRenderer / BinauralRenderer drive it
// (src/brsrenderer.h:186-205): per block  [old state] -> rotate_queues() if the
// queues were not empty -> set_filter() if the filter changed -> [new state].
static void test_filter_ring_rule()
{
  static float2 storage[64][64];
  static float2 empty_marker;
  for (int P : {1, 2, 3, 4, 7, 47})
  {
    LiteralQueues lit(P, &empty_marker);
    // ring[b] = index of the filter current in block b (-1: none yet) and its length
    std::vector<int> cur_filter(1, -1), cur_len(1, 0);   // block 0 = before the first block
    long long last_set = 0; bool has_set = false;
    std::vector<const float2*> prev_state(P, &empty_marker);
    for (long long b = 1; b <= 500; ++b)
    {
      // queues_empty before this block's rotate, by the block-index rule
      const bool rule_empty = P < 2 || !has_set || b >= last_set + P;
      CHECK(rule_empty == lit.queues_empty());
      // old state = the table as the previous block left it
      bool same_old = true;
      for (int p = 0; p < P; ++p) same_old = same_old && lit.ptrs[p] == prev_state[p];
      CHECK(same_old);
      if (!lit.queues_empty()) lit.rotate_queues();
      int f = cur_filter.back(), len = cur_len.back();
      const unsigned op = rnd() % 10;
      // bursts of switches in consecutive blocks, pauses longer than P blocks
      const bool change = (b % 97 < 40) ? op < 8 : (b % 97 < 40 + P + 3 ? false : op < 2);
      if (change)
      {
        f = int(rnd() % 64);
        len = 1 + int(rnd() % (P + 2));
        if (len > 64) len = 64;
        std::vector<const float2*> parts;
        for (int p = 0; p < len; ++p) parts.push_back(&storage[f][p]);
        lit.set_filter(parts);
        last_set = b; has_set = true;
      }
      cur_filter.push_back(f); cur_len.push_back(len);
      // new state: partition p = part p of the filter current in block b - p
      bool same_new = true;
      for (int p = 0; p < P; ++p)
      {
        const long long tau = b - p;
        const float2* want = &empty_marker;
        if (tau >= 1 && cur_filter[size_t(tau)] >= 0 && p < cur_len[size_t(tau)])
          want = &storage[cur_filter[size_t(tau)]][p];
        same_new = same_new && lit.ptrs[p] == want;
      }
      CHECK(same_new);
      prev_state = lit.ptrs;
    }
  }
}

static void test_dyn_split()
{
  for (int trial = 0; trial < 2000; ++trial)
  {
    int n[3];
    for (int& v : n) v = (rnd() % 3 == 0) ? 0 : int(rnd() % 200);
    const int Pmax = 1 + int(rnd() % 60);
    const int G = 3 + int(rnd() % 600);
    int next = 0;
    bool ok = true;
    for (int k = 0; k < 3; ++k)
    {
      int first = -1, count = -1;
      ssrk::dyn_split(n, Pmax, G, k, first, count);
      ok = ok && first == next;                       // contiguous, in group order
      ok = ok && ((n[k] == 0) == (count == 0));       // exactly the non-empty groups get CTAs
      const int terms = n[k] * Pmax;
      ok = ok && count <= std::max(1, (terms + ssrk::kDynMinTermsPerCta - 1) / ssrk::kDynMinTermsPerCta);
      if (count > 0)
      {
        // the kernel's chunking covers every term exactly once
        const int chunk = (terms + count - 1) / count;
        long covered = 0;
        for (int s = 0; s < count; ++s)
          covered += std::min(terms, (s + 1) * chunk) - std::min(terms, s * chunk);
        ok = ok && covered == terms;
      }
      next = first + count;
    }
    ok = ok && next <= G;
    CHECK(ok);
  }
}

// the split the docs quote for BASELINE config 5 on one GPU
static void test_cfg5_split()
{
  MacPlan plan;
  plan.MT = 4;
  const float4* h[4] = { nullptr, nullptr, nullptr, nullptr };
  for (int g = 0; g < 128; ++g)          // 512 outputs in groups of 4
  {
    plan.begin_group();
    for (int n = 0; n < 128; ++n)        // 128 sources x 47 partitions
      for (int p = 0; p < 47; ++p) plan.add_term(n, p, n, h);
    plan.end_group();
  }
  // bulk-copy staged MAC, one CTA per SM: one wave of 148 equal shares of the 770 048 terms
  // (5203 each); 124 of the 127 group boundaries fall inside a share (3 coincide with share
  // boundaries: 128 / 148 = 32 / 37) -> 272 partial spectra groups
  // (round 1 cut every group into 8: 1024 CTAs = 6.92 waves, 1024 partials)
  plan.make_work(148, Engine::kSplitCostTma);
  CHECK(plan.ncta == 148);
  CHECK(plan.npartials == 148 + 124);
  plan.make_work(148, Engine::kSplitCostTma, 7);   // SSR_B200_MAC_WAVES=7
  CHECK(plan.ncta == 1036);
  // sharded over 8 GPUs: 16 groups per GPU -> all 148 SMs (round 1: 16 x 9 = 144 CTAs)
  MacPlan shard;
  shard.MT = 4;
  for (int g = 0; g < 16; ++g)
  {
    shard.begin_group();
    for (int n = 0; n < 128; ++n)
      for (int p = 0; p < 47; ++p) shard.add_term(n, p, n, h);
    shard.end_group();
  }
  shard.make_work(148, Engine::kSplitCostTma);
  CHECK(shard.ncta == 148);
  CHECK(shard.npartials == 148 + 12);
  int longest = 0;
  for (int c = 0; c < shard.ncta; ++c)
  {
    int share = 0;
    for (int i = shard.cta_first.host[size_t(c)]; i < shard.cta_first.host[size_t(c) + 1]; ++i)
      share += shard.work.host[size_t(i)].y - shard.work.host[size_t(i)].x;
    longest = std::max(longest, share);
  }
  CHECK(longest == 651);                 // 96 256 terms / 148 (round 1: 6016 / 9 = 669)
}

// ControlQueue = apf::CommandQueue for the control setters (apf/apf/commandqueue.h:141-153,185-210):
// commands come out in the order they went in, a full ring is reported (not overwritten), and
// a producer thread racing the consumer loses nothing.
static void test_control_queue()
{
  auto q = std::unique_ptr<ControlQueue>(new ControlQueue);
  for (int i = 0; i < 10; ++i) CHECK(q->push(ControlQueue::Cmd{ControlQueue::REF_ORIENTATION, i, float(i), 0.f}));
  int seen = 0; bool ordered = true;
  q->drain([&](const ControlQueue::Cmd& c) { ordered = ordered && c.id == seen && c.a == float(seen); ++seen; });
  CHECK(seen == 10 && ordered);
  q->drain([&](const ControlQueue::Cmd&) { ++seen; });
  CHECK(seen == 10);   // empty now
  for (uint32_t i = 0; i < ControlQueue::kCapacity; ++i) CHECK(q->push(ControlQueue::Cmd{0, int(i), 0.f, 0.f}));
  CHECK(!q->push(ControlQueue::Cmd{0, -1, 0.f, 0.f}));   // full: reported after the retries, nothing overwritten
  seen = 0; ordered = true;
  q->drain([&](const ControlQueue::Cmd& c) { ordered = ordered && c.id == seen; ++seen; });
  CHECK(seen == int(ControlQueue::kCapacity) && ordered);
  // producer thread vs consumer: 200 000 commands through a 4096-slot ring
  const int total = 200000;
  std::thread producer([&] { for (int i = 0; i < total; ++i) while (!q->push(ControlQueue::Cmd{1, i, 0.f, 0.f})) {} });
  int next = 0; ordered = true;
  while (next < total) q->drain([&](const ControlQueue::Cmd& c) { ordered = ordered && c.id == next; ++next; });
  producer.join();
  CHECK(ordered && next == total);
}

int main()
{
  test_control_queue();
  test_cfg5_split();
  test_ptr_table();
  test_make_work();
  test_filter_ring_rule();
  test_dyn_split();
  std::printf("%s (%d assertions, %d failed)\n", failures ? "FAILED" : "All tests passed", checks, failures);
  return failures ? 1 : 0;
}
