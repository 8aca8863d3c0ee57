// pcs_seq.cpp -- host side of libpcs_seq behind the C ABI (include/pcs_seq.h):
// context, forest upload, sample groups, planner (tile grid + host multinomial),
// run / trace / injected-count drivers.  No CPU fallback: every compute entry
// point launches the sm_100a kernels of kernels.cu or fails.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pcs_seq.h"
#include "dev.hpp"
#include "flat.hpp"
#include "host_pool.hpp"
#include "kernels.hpp"
#include "plan_rng.hpp"

namespace {

thread_local std::string g_err;

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_OK(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
  } while (0)

void require(bool ok, const char* msg) {
  if (!ok) throw std::domain_error(msg);
}

// threads of the host phases (flattener, sample groups, planner): PCS_HOST_THREADS, else every core up to 32 --
// the phases are short and memory-bound, and every one of them starts its own threads: past a few dozen the
// thread start-up costs more than the extra cores give
unsigned host_threads() {
  const char* s = std::getenv("PCS_HOST_THREADS");
  if (s) {
    long v = std::atol(s);
    if (v >= 1 && v <= 1024) return static_cast<unsigned>(v);
  }
  return std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
}

void parallel_copy(void* dst, const void* src, size_t bytes) {
  const unsigned nt = std::max(1u, std::min(16u, host_threads()));
  if (bytes < (4u << 20) || nt == 1) {
    std::memcpy(dst, src, bytes);
    return;
  }
  const size_t chunk = (bytes / nt + 4095) & ~static_cast<size_t>(4095);
  pcs::HostPool::get().run((bytes + chunk - 1) / chunk, [=](size_t k) {
    const size_t off = k * chunk;
    std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, std::min(chunk, bytes - off));
  }, nt);
}

// PCS_DEVICE_INSTANCES=0: the flattener of an uploaded forest writes the instance table itself and sends it (the
// round-1 path, kept for A/B runs); default: the device builds it (flat.hpp: FlatForest::inst_deferred)
bool device_instances() {
  const char* e = std::getenv("PCS_DEVICE_INSTANCES");
  return !(e && std::string(e) == "0");
}

// The caller's output tables are usually fresh memory (a new R vector, a new numpy array): the host threads that
// copy the tables out of pinned memory take a page fault per 4 KB, 35 000 of them for C3's 145 MB -- as long as
// the tables' trip over the link.  Where the kernel offers transparent huge pages on request (THP mode `madvise`
// or `always`), ask for them: a fault then maps 2 MB.  Harmless where it does not apply (the call fails or is
// ignored); PCS_HUGE_OUT=0 skips it.
void advise_huge_pages(void* p, size_t bytes) {
  static const bool on = [] {
    const char* e = std::getenv("PCS_HUGE_OUT");
    return !(e && std::string(e) == "0");
  }();
  if (!on || !p || bytes < (4u << 20)) return;
  const uintptr_t page = 4096, lo = (reinterpret_cast<uintptr_t>(p) + page - 1) & ~(page - 1),
                  hi = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(page - 1);
  if (hi > lo) (void)madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
}

// device buffer from the device's stream-ordered memory pool (cudaMallocAsync): repeated
// upload / plan / free cycles reuse pooled memory instead of paying cudaMalloc each time
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t st = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count, cudaStream_t stream) {
    release();
    st = stream;
    n = count;
    if (count) CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), stream));
  }
  size_t bytes() const { return n * sizeof(T); }
  // through a pinned staging area: host threads copy `v` there (safe to free `v` afterwards), the DMA runs at
  // link speed and is left in flight; the caller synchronises before the staging area is written again
  template <class Vec>
  size_t upload_staged(const Vec& v, cudaStream_t stream, char* stage, size_t& stage_off) {
    alloc(v.size(), stream);
    if (!v.empty()) {
      parallel_copy(stage + stage_off, v.data(), bytes());
      CUDA_OK(cudaMemcpyAsync(p, stage + stage_off, bytes(), cudaMemcpyHostToDevice, stream));
      stage_off += (bytes() + 255) & ~static_cast<size_t>(255);
    }
    return bytes();
  }
  // synchronous w.r.t. the host buffer (pageable memory): safe to free `v` afterwards
  template <class Vec>
  size_t upload(const Vec& v, cudaStream_t stream) {
    alloc(v.size(), stream);
    if (!v.empty()) {
      CUDA_OK(cudaMemcpyAsync(p, v.data(), bytes(), cudaMemcpyHostToDevice, stream));
      CUDA_OK(cudaStreamSynchronize(stream));
    }
    return bytes();
  }
};

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// PCS_TIMING=1: host-side phase times on stderr
struct Lap {
  bool on = std::getenv("PCS_TIMING") != nullptr;
  double t = now_ms();
  void operator()(const char* what) {
    if (!on) return;
    double n = now_ms();
    std::fprintf(stderr, "[pcs host]    %-28s %8.2f ms\n", what, n - t);
    t = n;
  }
};

}  // namespace

struct pcs_ctx;
void release_ctx(pcs_ctx* cx);  // one user less; the last one of a destroyed context frees it

struct pcs_ctx {
  // Forests (and the flattened host views that borrowed a pinned block) keep their context alive: pcs_destroy on
  // a context that still has users only marks it, and the last user to go frees it -- a caller that tears things
  // down in the wrong order (an exception between the two frees, an interpreter shutting down) does not crash.
  std::atomic<int> users{0};
  std::atomic<bool> destroyed{false};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  void bind() const { CUDA_OK(cudaSetDevice(device)); }
  // pinned staging buffer for the result tables: D2H into pinned memory runs at link speed, the copy
  // into the caller's (pageable) buffers is spread over host threads
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  void* staging(size_t bytes) {
    if (bytes > pinned_bytes) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      pinned_bytes = 0;
      bytes += bytes / 8;  // forests of one study differ a little: do not re-pin for every one of them
      CUDA_OK(cudaHostAlloc(&pinned, bytes, cudaHostAllocPortable));  // every device of the process may DMA into it
      pinned_bytes = bytes;
    }
    return pinned;
  }
  // Pinned blocks lent to forests for their flat tables (pcs::FlatStore): the flattener writes the tables
  // where the H2D DMA reads them.  A freed forest hands its block back; the next upload takes it again, so a
  // session that uploads forest after forest pins memory once.
  struct PinnedBlock {
    void* p = nullptr;
    size_t bytes = 0;
  };
  std::mutex pool_mutex;
  std::vector<PinnedBlock> pool;
  PinnedBlock lend(size_t bytes) {
    {
      std::lock_guard<std::mutex> lock(pool_mutex);
      size_t best = pool.size();
      for (size_t i = 0; i < pool.size(); ++i)
        if (pool[i].bytes >= bytes && (best == pool.size() || pool[i].bytes < pool[best].bytes)) best = i;
      if (best != pool.size()) {
        PinnedBlock b = pool[best];
        pool.erase(pool.begin() + static_cast<std::ptrdiff_t>(best));
        return b;
      }
      if (!pool.empty()) {  // too small for this forest: do not keep it pinned next to the new one
        cudaFreeHost(pool.back().p);
        pool.pop_back();
      }
    }
    PinnedBlock b;
    b.bytes = bytes + bytes / 8;  // forests of one study differ a little: do not re-pin for every one of them
    bind();
    CUDA_OK(cudaHostAlloc(&b.p, b.bytes, cudaHostAllocPortable));
    return b;
  }
  void give_back(PinnedBlock b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lock(pool_mutex);
    pool.push_back(b);
    while (pool.size() > 2) {  // keep the two largest
      size_t smallest = 0;
      for (size_t i = 1; i < pool.size(); ++i)
        if (pool[i].bytes < pool[smallest].bytes) smallest = i;
      cudaFreeHost(pool[smallest].p);
      pool.erase(pool.begin() + static_cast<std::ptrdiff_t>(smallest));
    }
  }
  // run counters come back into pinned memory: a D2H copy into pageable memory would block the calling
  // thread until the stream has drained, and with it the host threads that copy finished tables out
  unsigned long long* pinned_counters = nullptr;  // [4]
  unsigned long long* counters_home() {
    if (!pinned_counters) CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&pinned_counters), 4 * sizeof(unsigned long long), cudaHostAllocDefault));
    return pinned_counters;
  }
  // tables other devices of this process add into (pcs_simulate_multi): plain cudaMalloc, kept between calls
  void* peer_tables = nullptr;
  size_t peer_tables_bytes = 0;
  void* peer_table(size_t bytes) {
    if (bytes > peer_tables_bytes) {
      if (peer_tables) cudaFree(peer_tables);
      peer_tables = nullptr;
      peer_tables_bytes = 0;
      CUDA_OK(cudaMalloc(&peer_tables, bytes));
      peer_tables_bytes = bytes;
    }
    return peer_tables;
  }
  // second stream: the tables of one sample travel to the host while the next sample is being sampled
  cudaStream_t copy_stream = nullptr;
  cudaStream_t copier() {
    if (!copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    return copy_stream;
  }
  // timing events of the per-sample sampler launches of a pipelined call (created on first use)
  std::vector<cudaEvent_t> time_ev;
  cudaEvent_t time_event(size_t i) {
    while (time_ev.size() <= i) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreate(&e));
      time_ev.push_back(e);
    }
    return time_ev[i];
  }
  // pinned slots the per-sample plan slices are uploaded from: slot k is written again only after its last DMA
  struct PlanSlot {
    char* p = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
    bool in_flight = false;
  };
  PlanSlot plan_slots[6];  // [4]: the haplotype lists of a forest's sample groups; [5]: inputs of its device-built instances
  PlanSlot& group_slot(size_t bytes) { return plan_slot(4, bytes); }
  PlanSlot& instance_slot(size_t bytes) { return plan_slot(5, bytes); }
  PlanSlot& plan_slot(size_t k, size_t bytes) {  // k in [0, 5]
    PlanSlot& sl = plan_slots[k];
    if (!sl.done) CUDA_OK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    if (sl.in_flight) {
      CUDA_OK(cudaEventSynchronize(sl.done));
      sl.in_flight = false;
    }
    if (bytes > sl.bytes) {
      if (sl.p) cudaFreeHost(sl.p);
      sl.p = nullptr;
      sl.bytes = 0;
      bytes += bytes / 4;
      CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&sl.p), bytes, cudaHostAllocDefault));
      sl.bytes = bytes;
    }
    return sl;
  }
  // events that mark the chunks of a pipelined device-to-host copy (created on first use)
  std::vector<cudaEvent_t> chunk_ev;
  cudaEvent_t chunk_event(size_t i) {
    while (chunk_ev.size() <= i) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      chunk_ev.push_back(e);
    }
    return chunk_ev[i];
  }
};


namespace {
struct GridCache;  // the tile geometry of the forest's last call (planner, below)
std::shared_ptr<GridCache> make_grid_cache();
}  // namespace

// host half of an uploaded forest: flattened view + output sample groups
struct HostForest {
  std::shared_ptr<GridCache> grid_cache = make_grid_cache();
  pcs::FlatForest flat;
  // pinned block lent by a context for flat.store; handed back when the last device copy of the forest dies
  pcs_ctx* lender = nullptr;
  pcs_ctx::PinnedBlock block;
  HostForest() = default;
  HostForest(const HostForest&) = delete;
  HostForest& operator=(const HostForest&) = delete;
  ~HostForest() {
    if (lender) {
      lender->give_back(block);
      release_ctx(lender);
    }
  }
  void borrow(pcs_ctx* cx, size_t bytes) {
    block = cx->lend(bytes);
    lender = cx;
    ++cx->users;
    flat.store.base = static_cast<char*>(block.p);
    flat.store.capacity = block.bytes;
  }
  // deferred instance table (flat.inst_deferred): host copies of what the device built, fetched on first use
  std::mutex inst_mutex;
  std::vector<pcs::Inst> inst_copy;
  std::vector<uint32_t> inst_off_copy;
  uint32_t n_groups = 0;
  std::vector<uint32_t> leaf_group;
  std::vector<uint32_t> group_cells;  // tumour cells per group
  std::vector<uint32_t> hap_list;
  std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> list_index;  // (group<<32|fragset) -> (off,n)
  // (kind, cell, allele) -> haplotype index, per chromosome; built on first use
  std::vector<std::vector<std::pair<uint64_t, uint32_t>>> lookup;

  static uint64_t list_key(uint32_t group, uint32_t fragset) { return (static_cast<uint64_t>(group) << 32) | fragset; }
  static uint64_t hap_key(uint8_t kind, uint32_t cell, uint16_t allele) {
    return (static_cast<uint64_t>(kind) << 56) | (static_cast<uint64_t>(cell) << 16) | allele;
  }

  uint64_t groups_version = 0;  // bumped by build_groups: device copies of hap_list go stale

  void build_groups(const uint32_t* lg, uint32_t ng) {
    ++groups_version;
    n_groups = ng;
    leaf_group.assign(lg, lg + flat.n_leaves);
    group_cells.assign(ng, 0);
    for (uint32_t l = 0; l < flat.n_leaves; ++l) {
      require(leaf_group[l] < ng, "leaf group out of range");
      ++group_cells[leaf_group[l]];
    }
    // haplotype lists per (group | normal kind, fragment set), concatenated in key order.  A fragment set
    // belongs to one chromosome, so a list holds haplotype indices of that chromosome in increasing order:
    // count, offsets in key order, fill -- the two passes over the haplotypes run per chromosome in parallel.
    const size_t n_fs = flat.fragsets.size(), n_keys = (static_cast<size_t>(ng) + 2) * n_fs;
    auto group_of = [&](const pcs::HapRec& r) {
      return r.kind == pcs::HAP_TUMOUR ? leaf_group[r.cell] : ng + (r.kind == pcs::HAP_NORMAL_PRENEO ? 1u : 0u);
    };
    std::vector<uint32_t> count(n_keys + 1, 0);  // key = group * n_fs + fragset: the order of list_key()
    auto per_chr = [&](const std::function<void(uint32_t)>& fn) {
      pcs::HostPool::get().run(flat.n_chr, [&](size_t c) { fn(static_cast<uint32_t>(c)); }, host_threads());
    };
    per_chr([&](uint32_t c) {
      for (const pcs::HapRec& r : flat.chr_haps[c]) ++count[static_cast<size_t>(group_of(r)) * n_fs + r.fragset + 1];
    });
    hap_list.clear();
    list_index.clear();
    for (size_t k = 0; k < n_keys; ++k) {
      if (count[k + 1])
        list_index[list_key(static_cast<uint32_t>(k / n_fs), static_cast<uint32_t>(k % n_fs))] = {count[k], count[k + 1]};
      count[k + 1] += count[k];
    }
    hap_list.resize(count[n_keys]);
    per_chr([&](uint32_t c) {  // count[key] = where the next haplotype of the list goes
      const auto& haps = flat.chr_haps[c];
      for (uint32_t h = 0; h < haps.size(); ++h) hap_list[count[static_cast<size_t>(group_of(haps[h])) * n_fs + haps[h].fragset]++] = h;
    });
  }

  void build_lookup() {
    if (!lookup.empty()) return;
    lookup.resize(flat.n_chr);
    for (uint32_t c = 0; c < flat.n_chr; ++c) {
      auto& v = lookup[c];
      const auto& haps = flat.chr_haps[c];
      v.reserve(haps.size());
      for (uint32_t h = 0; h < haps.size(); ++h) v.emplace_back(hap_key(haps[h].kind, haps[h].cell, haps[h].allele), h);
      std::sort(v.begin(), v.end());
    }
  }
};

struct pcs_flat {
  HostForest host;
  std::unique_ptr<char[]> block;  // lent to host.flat.store like the pinned block of an uploaded forest
};

struct pcs_forest {
  // declared first = destroyed last: the device buffers below are freed on the context's stream before the
  // context loses this user
  struct CtxUser {
    pcs_ctx* cx = nullptr;
    void hold(pcs_ctx* c) {
      cx = c;
      ++c->users;
    }
    ~CtxUser() {
      if (cx) release_ctx(cx);
    }
  } user;
  pcs_ctx* ctx = nullptr;
  std::shared_ptr<HostForest> host_ptr = std::make_shared<HostForest>();
  HostForest& host = *host_ptr;  // several devices may hold one flattened forest
  DevBuf<uint32_t> d_locus_pos, d_chr_locus_off, d_locus_inst_off, d_row_locus, d_hap_list;
  DevBuf<pcs::Inst> d_inst;
  uint64_t h2d_bytes = 0;

  // reference bases / alt strings: only for materialising reads (SAM output)
  std::vector<std::string> ref_chr;     // [n_chr] ASCII bases, empty = not loaded
  std::vector<uint32_t> alt_off;        // [n_mut+1]
  std::string alt_bytes;
  bool seq_dirty = true;
  DevBuf<uint8_t> d_ref, d_alt;
  DevBuf<uint64_t> d_chr_ref_off;
  DevBuf<uint32_t> d_alt_off;

  pcs::SeqData seq_data() {
    const pcs::FlatForest& F = host.flat;
    require(alt_off.size() == static_cast<size_t>(F.n_mut) + 1, "alt strings were not set (pcs_forest_set_alt)");
    if (seq_dirty) {
      ctx->bind();
      std::vector<uint64_t> off(F.n_chr + 1, 0);
      std::vector<uint8_t> all;
      for (uint32_t c = 0; c < F.n_chr; ++c) {
        off[c] = all.size();
        if (c < ref_chr.size()) all.insert(all.end(), ref_chr[c].begin(), ref_chr[c].end());
      }
      off[F.n_chr] = all.size();
      std::vector<uint8_t> alt(alt_bytes.begin(), alt_bytes.end());
      h2d_bytes += d_ref.upload(all, ctx->stream);
      h2d_bytes += d_chr_ref_off.upload(off, ctx->stream);
      h2d_bytes += d_alt.upload(alt, ctx->stream);
      h2d_bytes += d_alt_off.upload(alt_off, ctx->stream);
      seq_dirty = false;
    }
    return pcs::SeqData{d_ref.p, d_chr_ref_off.p, d_alt.p, d_alt_off.p};
  }
  bool has_reference(uint32_t c) const { return c < ref_chr.size() && !ref_chr[c].empty(); }

  pcs::DevForest dev() const {
    pcs::DevForest F;
    F.locus_pos = d_locus_pos.p;
    F.chr_locus_off = d_chr_locus_off.p;
    F.locus_inst_off = d_locus_inst_off.p;
    F.inst = reinterpret_cast<const uint4*>(d_inst.p);
    F.hap_list = d_hap_list.p;
    F.n_loci = static_cast<uint32_t>(host.flat.locus_pos.size());
    F.n_mut = host.flat.n_mut;
    return F;
  }

  void set_groups(const uint32_t* lg, uint32_t ng) {
    host.build_groups(lg, ng);
    upload_groups();
  }
  uint64_t uploaded_groups = 0;
  void upload_groups() {
    // through a pinned slot and left in flight, behind the flat tables' own DMA: the caller goes on to plan
    ctx->bind();
    cudaStream_t st = ctx->stream;
    const size_t bytes = host.hap_list.size() * sizeof(uint32_t);
    d_hap_list.alloc(host.hap_list.size(), st);
    if (bytes) {
      pcs_ctx::PlanSlot& slot = ctx->group_slot(bytes);
      parallel_copy(slot.p, host.hap_list.data(), bytes);
      CUDA_OK(cudaMemcpyAsync(d_hap_list.p, slot.p, bytes, cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaEventRecord(slot.done, st));
      slot.in_flight = true;
    }
    h2d_bytes += bytes;
    uploaded_groups = host.groups_version;
  }
  // a replica whose sibling changed the sample groups re-uploads the haplotype lists
  void sync_groups() {
    if (uploaded_groups != host.groups_version) upload_groups();
  }
  // a table in the lent (pinned) block is DMA'd from where the flattener wrote it and left in flight; anything
  // else goes through the context's pinned staging area
  template <class T>
  void upload_table(DevBuf<T>& dst, const pcs::Table<T>& src, std::vector<std::pair<DevBuf<T>*, const pcs::Table<T>*>>& staged) {
    cudaStream_t st = ctx->stream;
    if (src.empty() || host.flat.store.lent(src.data(), src.size() * sizeof(T))) {
      dst.alloc(src.size(), st);
      if (!src.empty()) CUDA_OK(cudaMemcpyAsync(dst.p, src.data(), dst.bytes(), cudaMemcpyHostToDevice, st));
      h2d_bytes += dst.bytes();
    } else {
      staged.emplace_back(&dst, &src);
    }
  }
  // the tables the flattener finishes first, sent while it is still numbering haplotypes (they are read in place
  // from the lent pinned block; tables outside it wait for upload_flat)
  bool loci_uploaded = false;
  void upload_loci_early() {
    const pcs::FlatForest& F = host.flat;
    if (!F.store.lent(F.locus_pos.data(), F.locus_pos.size() * 4) || !F.store.lent(F.row_locus.data(), F.row_locus.size() * 4)) return;
    ctx->bind();
    std::vector<std::pair<DevBuf<uint32_t>*, const pcs::Table<uint32_t>*>> none;
    h2d_bytes += d_chr_locus_off.upload(F.chr_locus_off, ctx->stream);
    upload_table(d_locus_pos, F.locus_pos, none);
    upload_table(d_row_locus, F.row_locus, none);
    loci_uploaded = none.empty();
  }
  void upload_flat() {
    ctx->bind();
    cudaStream_t st = ctx->stream;
    const pcs::FlatForest& F = host.flat;
    std::vector<std::pair<DevBuf<uint32_t>*, const pcs::Table<uint32_t>*>> staged32;
    std::vector<std::pair<DevBuf<pcs::Inst>*, const pcs::Table<pcs::Inst>*>> staged_inst;
    if (!loci_uploaded) {
      h2d_bytes += d_chr_locus_off.upload(F.chr_locus_off, st);  // tiny, pageable, synchronous: before the big ones
      upload_table(d_locus_pos, F.locus_pos, staged32);
      upload_table(d_row_locus, F.row_locus, staged32);
    }
    if (!F.inst_deferred) {
      upload_table(d_locus_inst_off, F.locus_inst_off, staged32);
      upload_table(d_inst, F.inst, staged_inst);
    }
    if (staged32.empty() && staged_inst.empty()) {
      if (F.inst_deferred) build_instances_on_device();
      return;
    }
    auto padded = [](size_t bytes) { return (bytes + 255) & ~static_cast<size_t>(255); };
    size_t total = 0;
    for (const auto& [dst, src] : staged32) total += padded(src->size() * 4);
    for (const auto& [dst, src] : staged_inst) total += padded(src->size() * sizeof(pcs::Inst));
    CUDA_OK(cudaStreamSynchronize(st));  // nothing in flight may still use the staging area
    char* stage = static_cast<char*>(ctx->staging(total));
    size_t off = 0;
    for (const auto& [dst, src] : staged32) h2d_bytes += dst->upload_staged(*src, st, stage, off);
    for (const auto& [dst, src] : staged_inst) h2d_bytes += dst->upload_staged(*src, st, stage, off);
    CUDA_OK(cudaStreamSynchronize(st));
    if (F.inst_deferred) build_instances_on_device();
  }
  // The instance table of a forest flattened with deferred instances: one mask byte and one length pair per row
  // go up (from the pinned block, where the flattener wrote them), the somatic placements and the chromosomes'
  // germline intervals through a pinned slot, and the device writes inst / locus_inst_off (kernels.cu:
  // build_instances_kernel).  Everything is queued on the context's stream; the temporaries are freed in stream order.
  void build_instances_on_device() {
    cudaStream_t st = ctx->stream;
    const pcs::FlatForest& F = host.flat;
    const uint32_t M = F.n_mut, L = static_cast<uint32_t>(F.locus_pos.size());
    d_inst.alloc(F.n_inst, st);
    d_locus_inst_off.alloc(static_cast<size_t>(L) + 1, st);
    DevBuf<uint8_t> d_mask;
    DevBuf<uint16_t> d_meta;
    DevBuf<pcs::Inst> d_som;
    DevBuf<uint32_t> d_small, d_scratch;
    std::vector<std::pair<DevBuf<uint8_t>*, const pcs::Table<uint8_t>*>> st8;
    std::vector<std::pair<DevBuf<uint16_t>*, const pcs::Table<uint16_t>*>> st16;
    const pcs::Table<uint8_t> mask_rows{const_cast<uint8_t*>(F.germ_mask.data()), M};  // the table has one spare byte
    upload_table(d_mask, mask_rows, st8);
    upload_table(d_meta, F.row_meta, st16);
    require(st8.empty() && st16.empty(), "internal: the deferred instance inputs are not in the pinned block");
    // small inputs: [chr_row_off (n_chr + 1) | germ_iv (8 n_chr)] and the somatic placements, through a pinned slot
    const size_t n_small = F.chr_row_off.size() + F.germ_iv.size();
    const size_t b_small = n_small * sizeof(uint32_t), b_som = F.som.size() * sizeof(pcs::Inst);
    const size_t pad_small = (b_small + 255) & ~static_cast<size_t>(255);
    pcs_ctx::PlanSlot& slot = ctx->instance_slot(pad_small + b_som + 256);
    std::memcpy(slot.p, F.chr_row_off.data(), F.chr_row_off.size() * sizeof(uint32_t));
    std::memcpy(slot.p + F.chr_row_off.size() * sizeof(uint32_t), F.germ_iv.data(), F.germ_iv.size() * sizeof(uint32_t));
    if (b_som) std::memcpy(slot.p + pad_small, F.som.data(), b_som);
    d_small.alloc(n_small, st);
    d_som.alloc(F.som.size(), st);
    CUDA_OK(cudaMemcpyAsync(d_small.p, slot.p, b_small, cudaMemcpyHostToDevice, st));
    if (b_som) CUDA_OK(cudaMemcpyAsync(d_som.p, slot.p + pad_small, b_som, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaEventRecord(slot.done, st));
    slot.in_flight = true;
    h2d_bytes += b_small + b_som;
    d_scratch.alloc(static_cast<size_t>(pcs::active_blocks(M)) + 1, st);
    CUDA_OK(pcs::launch_build_instances(st, d_mask.p, d_meta.p, d_row_locus.p, d_small.p, F.n_chr,
                                        d_small.p + F.chr_row_off.size(), reinterpret_cast<const uint4*>(d_som.p),
                                        static_cast<uint32_t>(F.som.size()), d_scratch.p, M, L,
                                        static_cast<uint32_t>(F.n_inst), reinterpret_cast<uint4*>(d_inst.p),
                                        d_locus_inst_off.p));
  }
  // host copies of the two device-built tables, for the few host paths that read instances (the rows carried by
  // sequenced cells: include_non_sequenced_mutations): fetched once, on first use
  void ensure_host_instances() {
    pcs::FlatForest& F = host.flat;
    if (!F.inst_deferred) return;
    std::lock_guard<std::mutex> lock(host.inst_mutex);
    if (!host.inst_copy.empty() || F.n_inst == 0) return;
    ctx->bind();
    host.inst_copy.resize(F.n_inst);
    host.inst_off_copy.resize(F.locus_pos.size() + 1);
    CUDA_OK(cudaMemcpyAsync(host.inst_copy.data(), d_inst.p, F.n_inst * sizeof(pcs::Inst), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(host.inst_off_copy.data(), d_locus_inst_off.p, host.inst_off_copy.size() * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    F.inst = pcs::Table<pcs::Inst>{host.inst_copy.data(), host.inst_copy.size()};
    F.locus_inst_off = pcs::Table<uint32_t>{host.inst_off_copy.data(), host.inst_off_copy.size()};
  }
  pcs_forest() = default;
  explicit pcs_forest(const pcs_forest& other, pcs_ctx* cx) : ctx(cx), host_ptr(other.host_ptr), host(*host_ptr) {}
};

// host half of a plan: tile grid of this shard, sampling tables, sequencer model
struct HostPlan {
  std::vector<pcs::Tile> tiles;         // this shard, staged in shared memory, heaviest first
  std::vector<pcs::Tile> tiles_global;  // this shard, too dense to stage
  std::vector<pcs::Tile> tiles_by_sample;   // `tiles` grouped by output sample (same order inside a sample)
  std::vector<uint32_t> sample_tile_off;    // [n_out_samples + 1]
  pcs::StageDims dims{};
  std::vector<pcs::Entry> entries;
  std::vector<uint32_t> entry_lo;      // low word of every entry's leaf scale (dev.hpp: exact_leaf)
  std::vector<uint32_t> insert_alias;  // [n][2] {keep threshold, alias column}
  pcs::SeqModel model{};
  pcs_plan_info info{};
};

struct pcs_plan {
  pcs_forest* forest = nullptr;
  HostPlan host;
  DevBuf<pcs::Tile> d_tiles, d_tiles_global;
  // the staged tiles once more, grouped by output sample (host-output runs launch sample by sample)
  DevBuf<pcs::Tile> d_tiles_by_sample;  // uploaded by the first run that needs it
  DevBuf<pcs::Entry> d_entries;
  DevBuf<uint32_t> d_entry_lo;
  DevBuf<uint32_t> d_insert_alias;
  DevBuf<uint32_t> d_depth, d_occ, d_cov;
  DevBuf<unsigned long long> d_counters;  // [0] reads placed [1] sum depth [2] sum occ [3] trace count
  uint64_t h2d_bytes = 0;
  const uint32_t* last_occ = nullptr;  // occurrence table of the last pcs_plan_run (device), for pcs_plan_result
  uint64_t launches_since_read = 0;    // kernels launched since the counters were last read (asynchronous steps)
};

// compact, column-major result of one call, resident on the device until fetched (include/pcs_seq.h)
struct pcs_result {
  pcs_forest* forest = nullptr;
  uint32_t n_rows = 0, n_samples = 0;
  DevBuf<uint32_t> d_rows, d_occ, d_cov;
  DevBuf<double> d_vaf;
  pcs_run_stats stats{};
};

namespace {

// Tile width.  PCS_TILE_BP overrides; otherwise 2^18 bp, narrowed for small jobs so that the grid still has
// a few thousand CTAs (148 SMs x 4 resident CTAs x several waves).  A function of the forest and the call's
// parameters only -- never of the number of GPUs -- so tile ids and results do not depend on sharding.
uint32_t tile_bp(uint64_t sequenced_bp_all_samples) {
  const char* s = std::getenv("PCS_TILE_BP");
  if (s) {
    long v = std::atol(s);
    if (v >= 1024) return static_cast<uint32_t>(v);
  }
  uint32_t w = 1u << 18;
  while (w > (1u << 13) && sequenced_bp_all_samples / w < 148ull * 32ull) w >>= 1;
  return w;
}

uint32_t stage_loci_cap() {
  const char* s = std::getenv("PCS_STAGE_LOCI");
  if (s) {
    long v = std::atol(s);
    if (v >= 0 && v <= 8192) return static_cast<uint32_t>(v);
  }
  return 1024;
}

// cumulative thresholds over the u32 range: pick the first i with draw <= thr[i]
std::vector<uint32_t> thresholds(const std::vector<double>& w) {
  double total = 0;
  for (double x : w) total += x;
  std::vector<uint32_t> thr(w.size());
  double cum = 0;
  for (size_t i = 0; i < w.size(); ++i) {
    cum += w[i];
    double b = std::floor(cum / total * 4294967296.0);
    if (i + 1 == w.size() || b >= 4294967296.0) b = 4294967296.0;
    thr[i] = b < 1.0 ? 0u : static_cast<uint32_t>(b - 1.0);
  }
  return thr;
}

void validate(const pcs_seq_params& P) {
  require(P.read_size >= 1 && P.read_size <= 65535, "read_size must be in [1, 65535]");
  require(P.coverage >= 0 && std::isfinite(P.coverage), "coverage must be a non-negative number");
  require(P.normal_only || (P.purity >= 0 && P.purity <= 1), "purity must belong to [0,1]");
  require(P.sequencer <= PCS_SEQ_BASIC_RANDOM, "Unsupported sequencer type");
  require(P.error_rate >= 0 && std::isfinite(P.error_rate), "The parameter \"error_rate\" must be a positive real number.");
  if (P.insert_size_mean > 0) {
    // get_bin_dist(): src/seq_simulation.cpp:431-451
    double q = static_cast<double>(P.insert_size_stddev) * P.insert_size_stddev / P.insert_size_mean;
    if (1 - q < 0)
      throw std::runtime_error("The insert size mean (" + std::to_string(P.insert_size_mean) +
                               ") must be greater than or equal to its variance (" +
                               std::to_string(P.insert_size_stddev) + "*" + std::to_string(P.insert_size_stddev) + "=" +
                               std::to_string(P.insert_size_stddev * P.insert_size_stddev) + ").");
  }
  uint32_t sc = P.shard_count ? P.shard_count : 1;
  require(P.shard_rank < sc, "shard_rank must be smaller than shard_count");
}

// Binomial(t, p) as selection thresholds over its support
void insert_table(uint32_t mean, uint32_t sd, std::vector<uint32_t>& alias, uint32_t& kmin, uint32_t& kmax) {
  double q = static_cast<double>(sd) * sd / mean;
  double p = 1 - q;
  if (!(p > 0)) {  // sd^2 == mean: Binomial(., 0) is always 0
    kmin = kmax = 0;
    alias.assign({4294967295u, 0u});
    return;
  }
  const double td = mean / p;
  require(td < 4294967296.0, "the insert size law Binomial(mean / p, p), p = 1 - sd^2 / mean, has too many trials");
  uint32_t t = static_cast<uint32_t>(td);
  // The law is tabulated where it is not negligible (pmf >= 1e-18).  That support lies within 9-13 standard
  // deviations of the mean, so only the window mean +- (16 sd + 128) is evaluated: t = mean / p itself can be
  // billions when sd^2 is close to the mean.
  const uint32_t w_lo = static_cast<uint32_t>(std::max(0.0, static_cast<double>(mean) - 16.0 * sd - 128.0));
  const uint32_t w_hi = static_cast<uint32_t>(std::min(static_cast<double>(t), static_cast<double>(mean) + 16.0 * sd + 128.0));
  std::vector<double> win(static_cast<size_t>(w_hi - w_lo) + 1, 0.0);
  auto pmf = [&](uint32_t k) -> double& { return win[k - w_lo]; };  // k in [w_lo, w_hi]
  if (p >= 1.0) {
    pmf(t) = 1.0;
  } else {
    for (uint32_t k = w_lo; k <= w_hi; ++k)
      pmf(k) = std::exp(std::lgamma(t + 1.0) - std::lgamma(k + 1.0) - std::lgamma(t - k + 1.0) +
                        k * std::log(p) + (t - k) * std::log1p(-p));
  }
  kmin = w_lo;
  kmax = w_hi;
  while (kmin < kmax && pmf(kmin) < 1e-18) ++kmin;
  while (kmax > kmin && pmf(kmax) < 1e-18) --kmax;
  // Walker / Vose alias table over the support [kmin, kmax]
  const uint32_t n = kmax - kmin + 1;
  std::vector<double> scaled(n);
  double total = 0;
  for (uint32_t i = 0; i < n; ++i) total += pmf(kmin + i);
  for (uint32_t i = 0; i < n; ++i) scaled[i] = pmf(kmin + i) / total * n;
  std::vector<uint32_t> small, large;
  for (uint32_t i = 0; i < n; ++i) (scaled[i] < 1.0 ? small : large).push_back(i);
  alias.assign(2 * static_cast<size_t>(n), 0);
  auto set = [&](uint32_t col, double keep, uint32_t other) {
    alias[2 * col] = static_cast<uint32_t>(std::min(4294967295.0, std::floor(keep * 4294967296.0)));
    alias[2 * col + 1] = other;
  };
  while (!small.empty() && !large.empty()) {
    const uint32_t s_ = small.back(), l_ = large.back();
    small.pop_back();
    set(s_, scaled[s_], l_);
    scaled[l_] -= 1.0 - scaled[s_];
    if (scaled[l_] < 1.0) {
      large.pop_back();
      small.push_back(l_);
    }
  }
  for (uint32_t i : large) set(i, 1.0, i);
  for (uint32_t i : small) set(i, 1.0, i);
}

struct OutSample {
  bool is_normal;
  uint32_t group;
};

// run fn(k) for k in [0, n) on the host threads; the first failure is rethrown as std::domain_error
template <class Fn>
void host_tasks(size_t n, Fn&& fn) {
  std::vector<std::string> errors(n);
  pcs::HostPool::get().run(n, [&](size_t k) {
    try {
      fn(k);
    } catch (const std::exception& e) {
      errors[k] = e.what();
      if (errors[k].empty()) errors[k] = "planning failed";
    }
  }, host_threads());
  for (const auto& e : errors)
    if (!e.empty()) throw std::domain_error(e);
}

// first element >= key of the sorted range [first, last), expected near `first`
const uint32_t* gallop(const uint32_t* first, const uint32_t* last, uint32_t key) {
  size_t step = 1;
  const uint32_t* lo = first;  // everything before lo is < key
  while (static_cast<size_t>(last - lo) > step && lo[step - 1] < key) {
    lo += step;
    step <<= 1;
  }
  return std::lower_bound(lo, std::min(lo + step, last), key);
}

// first element >= key of the sorted range [first, last), expected near `guess` (a pointer into the range): an
// exponential search in whichever direction the guess missed -- with a good guess, two or three probes that share
// a cache line instead of a binary search's dozen misses
const uint32_t* gallop_from(const uint32_t* first, const uint32_t* last, const uint32_t* guess, uint32_t key) {
  if (guess >= last) guess = last;
  if (guess < first) guess = first;
  if (guess == last || *guess >= key) {  // the answer is at or before the guess
    size_t step = 1;
    const uint32_t* hi = guess;          // everything from hi on is >= key
    while (static_cast<size_t>(hi - first) >= step && hi[-static_cast<ptrdiff_t>(step)] >= key) {
      hi -= step;
      step <<= 1;
    }
    const uint32_t* lo = static_cast<size_t>(hi - first) >= step ? hi - step : first;
    return std::lower_bound(lo, hi, key);
  }
  return gallop(guess, last, key);
}

// The part of a plan no output sample changes: parameters resolved, and per chromosome the tile GEOMETRY
// (window, staged loci, rows) -- pieces cut into windows of <= W bp and <= lcap staged loci.
struct PlanSetup {
  const HostForest* fo = nullptr;
  pcs_seq_params P{};
  std::vector<uint8_t> chr_mask;
  std::vector<OutSample> samples;
  uint32_t R = 0, mates = 1, kmin = 0, kmax = 0, W = 0, lcap = 0, shards = 1, normal_group = 0, dir_shift = 5;
  uint64_t reach = 0;
  bool paired = false;
  bool thin = false;  // single-end reads: staged tiles draw only the templates that can span a locus (dev.hpp: Tile)
  std::vector<uint32_t> insert_alias;
  struct ChrGrid {
    std::vector<pcs::Tile> tiles;         // chr, begin, len, l0, l1, r0, n_rows
    std::vector<uint32_t> piece_tile_off; // [pieces of the chromosome + 1]
  };
  std::shared_ptr<const std::vector<ChrGrid>> grid_ptr;  // kept by the forest: later calls with the same geometry reuse it
  const std::vector<ChrGrid>& grid_of() const { return *grid_ptr; }
  bool sequenced(uint32_t c) const { return chr_mask.empty() || chr_mask[c]; }
};

// the tile geometry a forest keeps for its next call: it depends on the forest and on these parameters only
struct GridKey {
  uint32_t R = 0, W = 0, lcap = 0;
  uint64_t reach = 0;
  bool thin = false;
  std::vector<uint8_t> chr_mask;
  bool operator==(const GridKey& o) const {
    return R == o.R && W == o.W && lcap == o.lcap && reach == o.reach && thin == o.thin && chr_mask == o.chr_mask;
  }
};
struct GridCache {
  std::mutex m;
  GridKey key;
  std::shared_ptr<const std::vector<PlanSetup::ChrGrid>> grid;
};
std::shared_ptr<GridCache> make_grid_cache() { return std::make_shared<GridCache>(); }

PlanSetup plan_setup(const HostForest& fo, const pcs_seq_params& P) {
  PlanSetup ps;
  ps.fo = &fo;
  ps.P = P;
  const pcs::FlatForest& F = fo.flat;
  if (P.chr_mask) ps.chr_mask.assign(P.chr_mask, P.chr_mask + F.n_chr);
  ps.P.chr_mask = nullptr;  // the caller's pointer is not kept
  ps.R = P.read_size;
  ps.paired = P.insert_size_mean > 0;
  ps.mates = ps.paired ? 2 : 1;
  const char* too_long = "the template (both reads and the insert) must be shorter than 2^30 bases";
  require(2ull * ps.R + P.insert_size_mean < (1ull << 30), too_long);  // before the insert law is tabulated
  if (ps.paired) insert_table(P.insert_size_mean, P.insert_size_stddev, ps.insert_alias, ps.kmin, ps.kmax);
  ps.reach = ps.paired ? 2ull * ps.R + ps.kmax : ps.R;
  // positions are 32-bit on the device: chromosome (< 2^31, checked by the flattener) + template must not wrap
  require(ps.reach <= (1ull << 30), too_long);
  if (!P.normal_only)
    for (uint32_t g = 0; g < fo.n_groups; ++g) ps.samples.push_back({false, g});
  if (P.normal_only || P.with_normal_sample) ps.samples.push_back({true, 0});
  ps.normal_group = fo.n_groups + (P.preneoplastic_in_normal ? 1u : 0u);
  uint64_t sequenced_bp = 0;
  for (uint32_t c = 0; c < F.n_chr; ++c)
    if (ps.sequenced(c)) sequenced_bp += F.chr_len[c];
  ps.W = tile_bp(sequenced_bp * ps.samples.size());
  ps.lcap = stage_loci_cap();
  {
    const char* e = std::getenv("PCS_THIN");  // PCS_THIN=0: draw every template (the round-1 sampler), for A/B runs
    ps.thin = !ps.paired && ps.lcap > 0 && !(e && std::string(e) == "0");
  }
  ps.shards = P.shard_count ? P.shard_count : 1;
  while ((((static_cast<uint64_t>(ps.W) + ps.reach) >> ps.dir_shift) + 1) > 4096) ++ps.dir_shift;

  GridCache& cache = *fo.grid_cache;
  GridKey key;
  key.R = ps.R; key.W = ps.W; key.lcap = ps.lcap; key.reach = ps.reach; key.thin = ps.thin; key.chr_mask = ps.chr_mask;
  {
    std::lock_guard<std::mutex> lock(cache.m);
    if (cache.grid && cache.key == key) {
      ps.grid_ptr = cache.grid;
      return ps;
    }
  }
  auto grid_new = std::make_shared<std::vector<PlanSetup::ChrGrid>>(F.n_chr);
  std::vector<PlanSetup::ChrGrid>& grid = *grid_new;
  Lap lap;
  host_tasks(F.n_chr, [&](size_t ci) {
    const uint32_t c = static_cast<uint32_t>(ci);
    PlanSetup::ChrGrid& g = grid[c];
    g.piece_tile_off.assign(F.chr_piece_off[c + 1] - F.chr_piece_off[c] + 1, 0);
    if (!ps.sequenced(c)) return;
    const uint32_t* lp = F.locus_pos.data();
    const uint32_t* c_lo = lp + F.chr_locus_off[c];
    const uint32_t* c_hi = lp + F.chr_locus_off[c + 1];
    const uint32_t* near = c_lo;  // pieces and windows come in increasing position: searches resume here
    const double density = static_cast<double>(c_hi - c_lo) / std::max<double>(1.0, F.chr_len[c]);
    for (uint32_t pi = F.chr_piece_off[c]; pi < F.chr_piece_off[c + 1]; ++pi) {
      const pcs::Piece& pc = F.pieces[pi];
      for (uint64_t b = pc.begin; b <= pc.end;) {
        pcs::Tile t{};
        t.chr = c;
        t.begin = static_cast<uint32_t>(b);
        near = gallop(near, c_hi, t.begin);
        t.l0 = static_cast<uint32_t>(near - lp);
        uint64_t len = std::min<uint64_t>(ps.W, pc.end - b + 1);
        for (;;) {  // shrink the window until its loci fit the staging capacity
          uint64_t last = std::min<uint64_t>(b + len + ps.reach, static_cast<uint64_t>(F.chr_len[c]) + 1);
          // the loci are spread roughly evenly: look for the window's end where the chromosome's mean density puts it
          const uint32_t* guess = near + static_cast<size_t>(static_cast<double>(last - b) * density);
          t.l1 = static_cast<uint32_t>(gallop_from(near, c_hi, guess, static_cast<uint32_t>(last)) - lp);
          if (t.l1 - t.l0 <= ps.lcap || len <= 2048) break;
          len = std::max<uint64_t>(2048, len / 2);
        }
        t.len = static_cast<uint32_t>(len);
        t.r0 = F.locus_first_row[t.l0];
        t.n_rows = F.locus_first_row[t.l1] - t.r0;
        t.tail_off = t.len;
        if (ps.thin && t.l1 - t.l0 <= std::min(ps.lcap, pcs::kMaxThinLoci) && t.n_rows <= 2 * ps.lcap) {
          // the offsets from which a read may run past the piece's end: the tail zone
          const uint64_t first_unsafe = static_cast<uint64_t>(pc.end) + 2 > static_cast<uint64_t>(ps.R) + b
                                            ? static_cast<uint64_t>(pc.end) + 2 - ps.R - b : 0;
          t.tail_off = static_cast<uint32_t>(std::min<uint64_t>(first_unsafe, len));
          t.thin = 1;  // u_len: second pass below
        }
        g.tiles.push_back(t);
        b += len;
      }
      g.piece_tile_off[pi - F.chr_piece_off[c] + 1] = static_cast<uint32_t>(g.tiles.size());
    }
  });
  lap("    geometry: windows");
  if (ps.thin) {
    // the useful offsets of every thinned tile (one scan of its loci): all tiles of the genome in equal chunks --
    // per chromosome, the longest one would be a tenth of the whole pass on one thread
    std::vector<std::pair<uint32_t, uint32_t>> chunks;  // (chromosome, first tile)
    constexpr uint32_t kChunk = 128;
    for (uint32_t c = 0; c < F.n_chr; ++c)
      for (uint32_t i = 0; i < grid[c].tiles.size(); i += kChunk) chunks.emplace_back(c, i);
    const uint32_t* lp = F.locus_pos.data();
    host_tasks(chunks.size(), [&](size_t k) {
      std::vector<pcs::Tile>& tiles = grid[chunks[k].first].tiles;
      const size_t i1 = std::min<size_t>(tiles.size(), chunks[k].second + kChunk);
      for (size_t i = chunks[k].second; i < i1; ++i) {
        pcs::Tile& t = tiles[i];
        if (!t.thin) continue;
        t.u_len = pcs::useful_offsets(lp + t.l0, t.l1 - t.l0, t.begin, t.len, t.tail_off, ps.R);
      }
    });
    lap("    geometry: useful offsets");
  }
  ps.grid_ptr = grid_new;
  {
    std::lock_guard<std::mutex> lock(cache.m);
    cache.key = std::move(key);
    cache.grid = grid_new;
  }
  return ps;
}

// tiles of one (output sample, chromosome): sampling entries per piece, template counts (multinomial over
// the tiles with the RNG stream of (seed, sample, chromosome)); entry_off is relative to `entries`
struct SampleChrPlan {
  std::vector<pcs::Entry> entries;
  std::vector<uint32_t> entry_lo;
  std::vector<pcs::Tile> tiles;
  std::vector<double> tile_w;        // share of the (sample, chromosome)'s templates every tile is drawn with
  std::vector<uint64_t> block_n;     // templates of every block of kTemplateBlock consecutive tiles
  uint64_t total_templates = 0;
};

// Templates per tile: a multinomial over the tiles of a (sample, chromosome), drawn in TWO LEVELS -- first over
// blocks of kTemplateBlock consecutive tiles (a chain of binomials, RNG stream of (seed, sample, chromosome)),
// then inside every block (stream of (seed, sample, chromosome, block)).  A multinomial split like this is the
// same law; the blocks are independent tasks, so the longest chromosome is no longer the planner's critical path.
constexpr size_t kTemplateBlock = 64;

void plan_block_templates(const PlanSetup& ps, uint32_t s, uint32_t c, SampleChrPlan& task, size_t b) {
  const size_t i0 = b * kTemplateBlock, i1 = std::min(task.tiles.size(), i0 + kTemplateBlock);
  pcs::PlanRng rng(static_cast<uint32_t>(ps.P.seed), 0x7116u, s, c, static_cast<uint32_t>(b));
  double wleft = 0;
  for (size_t i = i0; i < i1; ++i) wleft += task.tile_w[i];
  uint64_t left = task.block_n[b];
  for (size_t i = i0; i < i1 && left > 0; ++i) {
    const double p = (i + 1 == i1) ? 1.0 : std::min(1.0, std::max(0.0, task.tile_w[i] / wleft));
    const uint64_t k = pcs::binomial(rng, left, p);
    require(k <= 0xffffffffull, "too many templates in one tile; lower PCS_TILE_BP");
    pcs::Tile& t = task.tiles[i];
    t.n_templates = static_cast<uint32_t>(k);
    // of these, the templates whose read can span a locus: the start is uniform over the tile's offsets
    t.n_useful = t.thin ? static_cast<uint32_t>(pcs::binomial(rng, k, static_cast<double>(t.u_len) / t.len)) : t.n_templates;
    left -= k;
    wleft -= task.tile_w[i];
  }
}

void plan_sample_chr(const PlanSetup& ps, uint32_t s, uint32_t c, SampleChrPlan& task) {
  const HostForest& fo = *ps.fo;
  const pcs::FlatForest& F = fo.flat;
  const pcs_seq_params& P = ps.P;
  if (!ps.sequenced(c)) return;
  double purity = ps.samples[s].is_normal ? 0.0 : P.purity;
  const uint32_t nT = ps.samples[s].is_normal ? 0 : fo.group_cells[ps.samples[s].group];
  if (nT == 0) purity = 0.0;
  std::vector<pcs::Entry>& entries = task.entries;
  std::vector<pcs::Tile>& all = task.tiles;
  std::vector<double>& tile_w = task.tile_w;
  const PlanSetup::ChrGrid& g = ps.grid_of()[c];
  all.reserve(g.tiles.size());
  tile_w.reserve(g.tiles.size());
  // normal cells: every one carries each germline allele whole
  auto nit = fo.list_index.find(HostForest::list_key(ps.normal_group, F.full_fragset[c]));
  require(nit != fo.list_index.end(),
          P.preneoplastic_in_normal ? "preneoplastic_in_normal: the forest was uploaded without the normal cells that carry "
                                      "the pre-neoplastic mutations (pcs_cell_genomes_desc.n_normal_preneo = 0)"
                                    : "internal: normal haplotype list missing");
  const uint32_t n_normal_cells = P.preneoplastic_in_normal ? F.n_roots : 1;
  std::vector<double> w;
  std::vector<pcs::Entry> es;
  std::vector<uint32_t> es_n;  // haplotypes in each entry's list
  for (uint32_t pi = F.chr_piece_off[c]; pi < F.chr_piece_off[c + 1]; ++pi) {
    const pcs::Piece& pc = F.pieces[pi];
    w.clear();
    es.clear();
    es_n.clear();
    for (uint32_t k = 0; k < pc.cover_n; ++k) {
      const pcs::Cover& cv = F.covers[pc.cover_off + k];
      if (purity > 0) {
        auto it = fo.list_index.find(HostForest::list_key(ps.samples[s].group, cv.fragset));
        if (it != fo.list_index.end() && it->second.second > 0) {
          w.push_back(purity / nT * it->second.second);
          es.push_back(pcs::Entry{0u, 0u, it->second.first, cv.frag_end});
          es_n.push_back(it->second.second);
        }
      }
      if (purity < 1 && cv.fragset == F.full_fragset[c]) {
        w.push_back((1 - purity) / n_normal_cells * nit->second.second);
        es.push_back(pcs::Entry{0u, 0u, nit->second.first, cv.frag_end});
        es_n.push_back(nit->second.second);
      }
    }
    if (es.empty()) continue;
    double wsum = 0;
    for (double x : w) wsum += x;
    const std::vector<uint32_t> thr = thresholds(w);
    const uint32_t entry_off = static_cast<uint32_t>(entries.size());
    uint32_t n_kept = 0;
    {
      // entries whose share of the draw range is empty can never be picked: drop them
      uint64_t base = 0;
      for (size_t i = 0; i < es.size(); ++i) {
        if (static_cast<uint64_t>(thr[i]) + 1 <= base) continue;
        const uint64_t width = static_cast<uint64_t>(thr[i]) + 1 - base;
        es[i].thr = thr[i];
        // leaf = floor((u - base) * list_n / width), base = previous kept entry's thr + 1: uniform inside the entry
        uint32_t lo = 0;
        pcs::exact_leaf_scale(es_n[i], width, es[i].scale, lo);
        entries.push_back(es[i]);
        task.entry_lo.push_back(lo);
        ++n_kept;
        base = static_cast<uint64_t>(thr[i]) + 1;
      }
    }
    const uint32_t lp = pi - F.chr_piece_off[c];
    for (uint32_t ti = g.piece_tile_off[lp]; ti < g.piece_tile_off[lp + 1]; ++ti) {
      pcs::Tile t = g.tiles[ti];
      t.entry_off = entry_off;
      t.n_entries = n_kept;
      t.sample = s;
      if (n_kept > pcs::kMaxStagedEntries) t.thin = 0;  // goes to the global-memory kernel: drawn in full
      all.push_back(t);
      tile_w.push_back(wsum * t.len);
    }
  }
  // templates of this (sample, chromosome): N of them, multinomial over its tiles -- here over the BLOCKS of tiles
  // (the tiles inside a block are drawn by plan_block_templates, one independent task per block)
  const uint64_t N = static_cast<uint64_t>(std::llround(P.coverage * F.chr_len[c] / (static_cast<double>(ps.R) * ps.mates)));
  pcs::PlanRng rng(static_cast<uint32_t>(P.seed), 0x7115u, s, c, 0u);
  const size_t n_blocks = (all.size() + kTemplateBlock - 1) / kTemplateBlock;
  std::vector<double> block_w(n_blocks, 0.0);
  for (size_t i = 0; i < all.size(); ++i) block_w[i / kTemplateBlock] += tile_w[i];
  double wleft = 0;
  for (double x : block_w) wleft += x;
  task.block_n.assign(n_blocks, 0);
  uint64_t left = all.empty() ? 0 : N;
  task.total_templates = left;
  for (size_t b = 0; b < n_blocks && left > 0; ++b) {
    const double p = (b + 1 == n_blocks) ? 1.0 : std::min(1.0, std::max(0.0, block_w[b] / wleft));
    const uint64_t k = pcs::binomial(rng, left, p);
    task.block_n[b] = k;
    left -= k;
    wleft -= block_w[b];
  }
}

// tile indices by descending template count, ties by index: LSD radix sort of the counts (stable)
std::vector<uint32_t> heaviest_first(const std::vector<pcs::Tile>& all) {
  const size_t n = all.size();
  std::vector<uint32_t> order(n), tmp(n);
  for (size_t i = 0; i < n; ++i) order[i] = static_cast<uint32_t>(i);
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = 11 * pass;
    uint32_t cnt[2049] = {0};
    for (size_t i = 0; i < n; ++i) ++cnt[((~all[order[i]].n_templates >> shift) & 2047u) + 1];
    for (int d = 0; d < 2048; ++d) cnt[d + 1] += cnt[d];
    for (size_t i = 0; i < n; ++i) tmp[cnt[(~all[order[i]].n_templates >> shift) & 2047u]++] = order[i];
    order.swap(tmp);
  }
  return order;
}

void set_model(HostPlan& pl, const PlanSetup& ps) {
  pcs::SeqModel& M = pl.model;
  M.insert_alias = nullptr;
  M.read_size = ps.R;
  M.paired = ps.paired ? 1 : 0;
  M.sequencer = ps.P.sequencer;
  M.err_thr = static_cast<uint32_t>(std::min(4294967295.0, std::floor(ps.P.error_rate * 4294967296.0)));
  if (ps.P.sequencer == PCS_SEQ_BASIC_RANDOM) {
    // random quality: a bound no base's error probability exceeds -- the ramp is <= 1.5 and the quality deviate
    // <= sqrt(-2 ln(2^-33)) = 6.77 (Box-Muller of a 32-bit uniform), taken as 6.9 against float rounding; sigma =
    // 0.5 (kernels.cu: kQualSigma).  A test word at or above it cannot be an error, so the samplers skip the
    // quality model for it (kernels.cu: pair_error_codes); 2 % of slack covers u01()'s float rounding of the word.
    const double bound = ps.P.error_rate * 1.5 * std::exp(0.5 * 6.9 - 0.125) * 1.02 + 1e-7;
    M.err_thr = bound >= 1.0 ? 4294967295u : static_cast<uint32_t>(std::min(4294967295.0, std::ceil(bound * 4294967296.0)));
  }
  M.error_rate = static_cast<float>(ps.P.error_rate);
  M.insert_n = static_cast<uint32_t>(ps.insert_alias.size() / 2);
  M.insert_min = ps.kmin;
  M.seed = static_cast<uint32_t>(ps.P.seed);
  M.reach = static_cast<uint32_t>(ps.reach);
  M.dir_shift = ps.dir_shift;
}

// what a tile costs the sampler, in warp instructions per 32 templates (profiles/: Philox + draw + probe ~35
// for every read, queue + walk + counting ~160 x the fraction of reads that reach a locus): shards are balanced
// on this, not on templates -- tiles differ in locus density
uint64_t tile_cost(const pcs::Tile& t, const PlanSetup& ps) {
  static const bool by_templates = [] {  // PCS_BALANCE=templates: the round-1 rule, for A/B measurements
    const char* e = std::getenv("PCS_BALANCE");
    return e && std::string(e) == "templates";
  }();
  if (by_templates) return t.n_templates;
  if (t.thin) return static_cast<uint64_t>(t.n_useful) * 130u + 2000u;  // every drawn read reaches a locus
  const double span = static_cast<double>(t.len) + static_cast<double>(ps.reach);
  const double hit = std::min(1.0, static_cast<double>(t.l1 - t.l0) * ps.R * ps.mates / span / ps.mates);
  return static_cast<uint64_t>(static_cast<double>(t.n_templates) * ps.mates * (35.0 + 160.0 * hit));
}

// The plan of the output samples [s0, s1) of the job: the slice of the whole job's plan that belongs to them.
// Tile ids are positions in the (sample, chromosome, position) order of the WHOLE job, so `id_base` = tiles of
// the samples before s0 (info.n_tiles_total of their plans); template counts come from the RNG stream of
// (seed, sample, chromosome).  Results do not depend on how the job is cut into plans or shards.
// `all_shards`: one plan per shard of ps.shards (one process driving several GPUs plans the job once); else only
// the plan of shard ps.P.shard_rank.
std::vector<HostPlan> make_host_plans(const PlanSetup& ps, uint32_t s0, uint32_t s1, uint32_t id_base, bool all_shards) {
  std::vector<HostPlan> plans(all_shards ? ps.shards : 1u);
  HostPlan& pl = plans[0];
  Lap lap;
  const HostForest& fo = *ps.fo;
  const pcs::FlatForest& F = fo.flat;
  // every (output sample, chromosome) is an independent task with its own RNG stream
  std::vector<SampleChrPlan> per(static_cast<size_t>(s1 - s0) * F.n_chr);
  host_tasks(per.size(), [&](size_t k) {
    plan_sample_chr(ps, s0 + static_cast<uint32_t>(k / F.n_chr), static_cast<uint32_t>(k % F.n_chr), per[k]);
  });
  {
    std::vector<std::pair<uint32_t, uint32_t>> blocks;  // (task, block)
    for (size_t k = 0; k < per.size(); ++k)
      for (size_t b = 0; b < per[k].block_n.size(); ++b) blocks.emplace_back(static_cast<uint32_t>(k), static_cast<uint32_t>(b));
    host_tasks(blocks.size(), [&](size_t q) {
      const uint32_t k = blocks[q].first;
      plan_block_templates(ps, s0 + k / F.n_chr, k % F.n_chr, per[k], blocks[q].second);
    });
  }
  lap("  plan: entries + templates");
  // merged in (sample, chromosome) order: the position in this order is the tile id
  std::vector<pcs::Entry>& entries = pl.entries;
  std::vector<pcs::Tile> all;
  {
    size_t n_tiles = 0, n_entries = 0;
    for (const auto& t : per) {
      n_tiles += t.tiles.size();
      n_entries += t.entries.size();
    }
    all.reserve(n_tiles);
    entries.reserve(n_entries);
    pl.entry_lo.reserve(n_entries);
  }
  uint64_t total_templates = 0;
  for (auto& task : per) {
    const uint32_t base = static_cast<uint32_t>(entries.size());
    entries.insert(entries.end(), task.entries.begin(), task.entries.end());
    pl.entry_lo.insert(pl.entry_lo.end(), task.entry_lo.begin(), task.entry_lo.end());
    for (auto& t : task.tiles) {
      t.entry_off += base;
      t.id = id_base + static_cast<uint32_t>(all.size());
      all.push_back(t);
    }
    total_templates += task.total_templates;
  }
  lap("  plan: merge");
  // shard: longest-processing-time greedy on the tiles' cost, heaviest first; ties by tile id => deterministic
  const std::vector<uint32_t> order = heaviest_first(all);
  // tiles whose loci / instances / rows fit the staging capacity go to the staged kernel
  const uint32_t lcap = ps.lcap, dir_shift = ps.dir_shift, shards = ps.shards;
  const uint64_t reach = ps.reach;
  auto stageable = [&](const pcs::Tile& t) {
    return lcap > 0 && t.l1 - t.l0 <= lcap && t.n_rows <= 2 * lcap && t.n_entries <= pcs::kMaxStagedEntries;
  };
  auto take = [&](HostPlan& dst, const pcs::Tile& t) {
    if (stageable(t)) {
      dst.tiles.push_back(t);
      dst.dims.max_loci = std::max(dst.dims.max_loci, t.l1 - t.l0);
      dst.dims.max_rows = std::max(dst.dims.max_rows, t.n_rows);
      dst.dims.max_buckets = std::max<uint32_t>(dst.dims.max_buckets, static_cast<uint32_t>(((t.len + reach) >> dir_shift) + 1));
      // a thinned tile keeps a 2-byte slot per 64 useful offsets where the others keep their 8-byte buckets
      if (t.thin) dst.dims.max_buckets = std::max<uint32_t>(dst.dims.max_buckets, (t.u_len >> 8) + 2);
    } else {
      dst.tiles_global.push_back(t);
    }
  };
  for (auto& q : plans) q.tiles.reserve(all.size() / shards + 16);
  std::vector<uint64_t> templates(shards, 0);
  if (shards == 1) {
    for (uint32_t i : order)
      if (all[i].n_templates) take(pl, all[i]);
    templates[0] = total_templates;
  } else {
    std::vector<uint64_t> load(shards, 0);
    for (uint32_t i : order) {
      if (!all[i].n_templates) continue;
      uint32_t best = 0;
      for (uint32_t r = 1; r < shards; ++r)
        if (load[r] < load[best]) best = r;
      load[best] += tile_cost(all[i], ps);
      templates[best] += all[i].n_templates;
      if (all_shards) take(plans[best], all[i]);
      else if (best == ps.P.shard_rank) take(pl, all[i]);
    }
  }
  lap("  plan: order + shard");
  const size_t S = ps.samples.size();
  for (size_t q = 0; q < plans.size(); ++q) {
    HostPlan& d = plans[q];
    if (q != 0) {  // every shard's tiles index the same sampling entries
      d.entries = pl.entries;
      d.entry_lo = pl.entry_lo;
    }
    // the staged tiles once more, grouped by output sample, heaviest first inside a sample: host-output runs
    // launch sample by sample
    d.sample_tile_off.assign(S + 1, 0);
    for (const auto& t : d.tiles) ++d.sample_tile_off[t.sample + 1];
    for (size_t i = 0; i < S; ++i) d.sample_tile_off[i + 1] += d.sample_tile_off[i];
    {
      std::vector<uint32_t> at(d.sample_tile_off.begin(), d.sample_tile_off.end() - 1);
      d.tiles_by_sample.resize(d.tiles.size());
      for (const auto& t : d.tiles) d.tiles_by_sample[at[t.sample]++] = t;
    }
    // round the capacities so that plans of similar forests share one shared-memory footprint
    d.dims.max_loci = (d.dims.max_loci + 63) & ~63u;
    d.dims.max_rows = (d.dims.max_rows + 63) & ~63u;
    d.dims.max_buckets = (d.dims.max_buckets + 63) & ~63u;
    d.insert_alias = ps.insert_alias;
    set_model(d, ps);
    d.info.n_out_samples = static_cast<uint32_t>(ps.samples.size());
    d.info.n_mut = F.n_mut;
    d.info.n_loci = static_cast<uint32_t>(F.locus_pos.size());
    d.info.n_tiles = d.tiles.size() + d.tiles_global.size();
    d.info.n_tiles_total = all.size();
    d.info.n_templates = templates[all_shards ? q : (shards == 1 ? 0 : ps.P.shard_rank)];
    d.info.n_templates_total = total_templates;
    d.info.reads_per_template = ps.mates;
    d.info.read_size = ps.R;
  }
  return plans;
}

HostPlan make_host_plan(const PlanSetup& ps, uint32_t s0, uint32_t s1, uint32_t id_base) {
  return std::move(make_host_plans(ps, s0, s1, id_base, false)[0]);
}

HostPlan make_host_plan(const HostForest& fo, const pcs_seq_params& P) {
  Lap lap;
  const PlanSetup ps = plan_setup(fo, P);
  lap("  plan: tile geometry");
  return make_host_plan(ps, 0, static_cast<uint32_t>(ps.samples.size()), 0);
}

void upload_plan(pcs_plan& pl) {
  pcs_forest& fo = *pl.forest;
  fo.ctx->bind();
  cudaStream_t st = fo.ctx->stream;
  pl.h2d_bytes += pl.d_tiles.upload(pl.host.tiles, st);
  pl.h2d_bytes += pl.d_tiles_global.upload(pl.host.tiles_global, st);
  pl.d_tiles_by_sample.release();
  pl.h2d_bytes += pl.d_entries.upload(pl.host.entries, st);
  pl.h2d_bytes += pl.d_entry_lo.upload(pl.host.entry_lo, st);
  pl.h2d_bytes += pl.d_insert_alias.upload(pl.host.insert_alias, st);
  pl.host.model.insert_alias = pl.d_insert_alias.p;
  const size_t S = pl.host.info.n_out_samples;
  pl.d_depth.alloc(S * pl.host.info.n_loci, st);
  pl.d_occ.alloc(S * pl.host.info.n_mut, st);
  pl.d_cov.alloc(S * pl.host.info.n_mut, st);
  pl.d_counters.alloc(4, st);
}

void run_plan(pcs_plan& pl, int flags, uint32_t* occ, uint32_t* cov, pcs_run_stats* stats) {
  pcs_forest& fo = *pl.forest;
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  const bool dev_out = (flags & PCS_RUN_DEVICE_OUTPUT) != 0;
  const bool checksums = (flags & PCS_RUN_NO_CHECKSUMS) == 0;
  const bool async = dev_out && (flags & PCS_RUN_ASYNC) != 0;  // nothing comes back to the host: no need to wait
  const size_t S = pl.host.info.n_out_samples, M = pl.host.info.n_mut, L = pl.host.info.n_loci;
  uint32_t* d_occ = dev_out ? occ : pl.d_occ.p;
  uint32_t* d_cov = dev_out ? cov : pl.d_cov.p;
  require(S * M == 0 || (occ && cov), "occurrences/coverage output pointers are NULL");
  pl.last_occ = d_occ;
  const double t0 = now_ms();
  uint64_t launches = 0;

  CUDA_OK(cudaEventRecord(cx.ev[0], st));
  if (S * L != 0) CUDA_OK(cudaMemsetAsync(pl.d_depth.p, 0, S * L * sizeof(uint32_t), st));
  if (S * M != 0) CUDA_OK(cudaMemsetAsync(d_occ, 0, S * M * sizeof(uint32_t), st));
  if (!async) CUDA_OK(cudaMemsetAsync(pl.d_counters.p, 0, 4 * sizeof(unsigned long long), st));  // async: they accumulate
  CUDA_OK(cudaEventRecord(cx.ev[1], st));
  const pcs::DevForest DF = fo.dev();
  uint64_t d2h = 0;
  const size_t table_bytes = S * M * sizeof(uint32_t);
  // tables back to the caller's (pageable) buffers: the DMA lands in pinned memory chunk by chunk, and host
  // threads copy chunk k out while chunk k+1 is still on the link
  struct Chunk {
    char* dst;
    const char* src;
    size_t bytes;
  };
  std::vector<Chunk> chunks;
  // host output, several samples, everything staged: sample by sample, so that the tables of sample s cross
  // the link (second stream) while sample s+1 is being sampled.  Same tiles, same counters, same tables.
  const bool by_sample = !dev_out && S > 1 && M != 0 && pl.host.tiles_global.empty() && !pl.host.tiles.empty() &&
                         std::getenv("PCS_NO_SPLIT") == nullptr;
  if (by_sample) {
    if (pl.d_tiles_by_sample.p == nullptr) {
      pl.h2d_bytes += pl.d_tiles_by_sample.upload(pl.host.tiles_by_sample, st);
      CUDA_OK(cudaEventRecord(cx.ev[1], st));  // the upload above is not kernel time
    }
    cudaStream_t cs = cx.copier();
    char* stage = static_cast<char*>(cx.staging(2 * table_bytes));
    const size_t row_bytes = M * sizeof(uint32_t);
    for (size_t smp = 0; smp < S; ++smp) {
      const uint32_t t0 = pl.host.sample_tile_off[smp], t1 = pl.host.sample_tile_off[smp + 1];
      CUDA_OK(pcs::launch_sample_tiles_staged(st, pl.d_tiles_by_sample.p + t0, t1 - t0, pl.d_entries.p, pl.d_entry_lo.p, DF, pl.host.model,
                                              pl.host.dims, pl.d_depth.p, d_occ, pl.d_counters.p));
      CUDA_OK(pcs::launch_finalize(st, pl.d_depth.p + smp * L, fo.d_row_locus.p, 1u, static_cast<uint32_t>(L),
                                   static_cast<uint32_t>(M), d_cov + smp * M));
      launches += (t1 > t0 ? 1 : 0) + 1;
      cudaEvent_t done = cx.chunk_event(2 * S + smp);
      CUDA_OK(cudaEventRecord(done, st));
      CUDA_OK(cudaStreamWaitEvent(cs, done, 0));
      const char* dev_tbl[2] = {reinterpret_cast<const char*>(d_occ + smp * M), reinterpret_cast<const char*>(d_cov + smp * M)};
      char* host_tbl[2] = {reinterpret_cast<char*>(occ + smp * M), reinterpret_cast<char*>(cov + smp * M)};
      for (int t = 0; t < 2; ++t) {
        char* sp = stage + (2 * smp + t) * row_bytes;
        CUDA_OK(cudaMemcpyAsync(sp, dev_tbl[t], row_bytes, cudaMemcpyDeviceToHost, cs));
        CUDA_OK(cudaEventRecord(cx.chunk_event(chunks.size()), cs));
        chunks.push_back({host_tbl[t], sp, row_bytes});
      }
    }
    d2h += 2 * table_bytes;
    CUDA_OK(cudaEventRecord(cx.ev[2], st));
  } else {
    CUDA_OK(pcs::launch_sample_tiles_staged(st, pl.d_tiles.p, static_cast<uint32_t>(pl.host.tiles.size()), pl.d_entries.p,
                                            pl.d_entry_lo.p, DF, pl.host.model, pl.host.dims, pl.d_depth.p, d_occ, pl.d_counters.p));
    CUDA_OK(pcs::launch_sample_tiles_global(st, pl.d_tiles_global.p, static_cast<uint32_t>(pl.host.tiles_global.size()),
                                            pl.d_entries.p, pl.d_entry_lo.p, DF, pl.host.model, pl.d_depth.p, d_occ, pl.d_counters.p));
    launches += (pl.host.tiles.empty() ? 0 : 1) + (pl.host.tiles_global.empty() ? 0 : 1);
    CUDA_OK(cudaEventRecord(cx.ev[2], st));
    CUDA_OK(pcs::launch_finalize(st, pl.d_depth.p, fo.d_row_locus.p, static_cast<uint32_t>(S), static_cast<uint32_t>(L),
                                 static_cast<uint32_t>(M), d_cov));
    launches += S * M != 0 ? 1 : 0;
  }
  if (checksums) {
    CUDA_OK(pcs::launch_sum_u32(st, pl.d_depth.p, S * L, pl.d_counters.p + 1));
    CUDA_OK(pcs::launch_sum_u32(st, d_occ, S * M, pl.d_counters.p + 2));
    launches += (S * M != 0 ? 1 : 0) + (S * L != 0 ? 1 : 0);
  }
  CUDA_OK(cudaEventRecord(cx.ev[3], st));
  pl.launches_since_read += launches;
  if (async) return;
  if (!by_sample && !dev_out && table_bytes != 0) {
    char* stage = static_cast<char*>(cx.staging(2 * table_bytes));
    const size_t cb = std::max<size_t>(4u << 20, ((2 * table_bytes / 16) + 4095) & ~static_cast<size_t>(4095));
    const char* dev_tbl[2] = {reinterpret_cast<const char*>(d_occ), reinterpret_cast<const char*>(d_cov)};
    char* host_tbl[2] = {reinterpret_cast<char*>(occ), reinterpret_cast<char*>(cov)};
    for (int t = 0; t < 2; ++t)
      for (size_t off = 0; off < table_bytes; off += cb) {
        const size_t nb = std::min(cb, table_bytes - off);
        char* sp = stage + t * table_bytes + off;
        CUDA_OK(cudaMemcpyAsync(sp, dev_tbl[t] + off, nb, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaEventRecord(cx.chunk_event(chunks.size()), st));
        chunks.push_back({host_tbl[t] + off, sp, nb});
      }
    d2h += 2 * table_bytes;
  }
  // asynchronous (pinned destination): the host threads below start copying finished chunks out while the
  // sampler is still running
  unsigned long long* counters = cx.counters_home();
  CUDA_OK(cudaMemcpyAsync(counters, pl.d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (!chunks.empty()) {
    const unsigned nt = std::max(1u, std::min(16u, host_threads()));
    std::vector<cudaError_t> werr(nt, cudaSuccess);
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nt; ++w)
      th.emplace_back([&, w] {
        cudaSetDevice(cx.device);
        for (size_t k = 0; k < chunks.size(); ++k) {
          const cudaError_t e = cudaEventSynchronize(cx.chunk_ev[k]);
          if (e != cudaSuccess) {
            werr[w] = e;
            return;
          }
          const size_t lo = (chunks[k].bytes * w / nt) & ~static_cast<size_t>(63);
          const size_t hi = w + 1 == nt ? chunks[k].bytes : (chunks[k].bytes * (w + 1) / nt) & ~static_cast<size_t>(63);
          if (hi > lo) std::memcpy(chunks[k].dst + lo, chunks[k].src + lo, hi - lo);
        }
      });
    for (auto& t : th) t.join();
    for (cudaError_t e : werr) CUDA_OK(e);
  }
  CUDA_OK(cudaStreamSynchronize(st));
  if (by_sample) CUDA_OK(cudaStreamSynchronize(cx.copy_stream));
  d2h += 4 * sizeof(unsigned long long);
  if (std::getenv("PCS_TIMING")) std::fprintf(stderr, "[pcs host]    %-28s %8.2f ms\n", "run (kernels + D2H)", now_ms() - t0);
  if (stats) {
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, cx.ev[1], cx.ev[2]));
    stats->kernel_ms = ms;
    stats->total_ms = now_ms() - t0;
    stats->kernel_launches = launches;
    stats->n_templates = pl.host.info.n_templates;
    stats->n_reads = counters[0];
    stats->sum_depth = counters[1];
    stats->sum_occurrences = counters[2];
    stats->h2d_bytes = 0;
    stats->d2h_bytes = d2h;
  }
}

// the sampler of this shard, adding into caller-provided tables (possibly peer memory): no zeroing, no finalize
void accumulate_plan(pcs_plan& pl, uint32_t* d_depth, uint32_t* d_occ, pcs_run_stats* stats) {
  pcs_forest& fo = *pl.forest;
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  const size_t S = pl.host.info.n_out_samples, M = pl.host.info.n_mut, L = pl.host.info.n_loci;
  require(S * M == 0 || (d_depth && d_occ), "depth/occurrences table pointers are NULL");
  (void)L;
  const double t0 = now_ms();
  // stats == NULL: asynchronous -- nothing is read back, the call returns once the kernels are queued and the
  // counters keep accumulating until pcs_plan_counters() reads them
  if (stats) CUDA_OK(cudaMemsetAsync(pl.d_counters.p, 0, 4 * sizeof(unsigned long long), st));
  CUDA_OK(cudaEventRecord(cx.ev[1], st));
  const pcs::DevForest DF = fo.dev();
  CUDA_OK(pcs::launch_sample_tiles_staged(st, pl.d_tiles.p, static_cast<uint32_t>(pl.host.tiles.size()), pl.d_entries.p,
                                          pl.d_entry_lo.p, DF, pl.host.model, pl.host.dims, d_depth, d_occ, pl.d_counters.p));
  CUDA_OK(pcs::launch_sample_tiles_global(st, pl.d_tiles_global.p, static_cast<uint32_t>(pl.host.tiles_global.size()),
                                          pl.d_entries.p, pl.d_entry_lo.p, DF, pl.host.model, d_depth, d_occ, pl.d_counters.p));
  CUDA_OK(cudaEventRecord(cx.ev[2], st));
  pl.launches_since_read += (pl.host.tiles.empty() ? 0 : 1) + (pl.host.tiles_global.empty() ? 0 : 1);
  if (!stats) return;
  unsigned long long counters[4] = {0, 0, 0, 0};
  CUDA_OK(cudaMemcpyAsync(counters, pl.d_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  if (stats) {
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, cx.ev[1], cx.ev[2]));
    *stats = pcs_run_stats{};
    stats->kernel_ms = ms;
    stats->total_ms = now_ms() - t0;
    stats->kernel_launches = (pl.host.tiles.empty() ? 0 : 1) + (pl.host.tiles_global.empty() ? 0 : 1);
    stats->n_templates = pl.host.info.n_templates;
    stats->n_reads = counters[0];
    stats->d2h_bytes = sizeof(counters);
  }
}

// coverage[s][row] = depth[s][locus(row)] and the table checksums, once every shard has accumulated
void finalize_tables(pcs_plan& pl, const uint32_t* d_depth, const uint32_t* d_occ, uint32_t* d_cov, pcs_run_stats* stats) {
  pcs_forest& fo = *pl.forest;
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  const size_t S = pl.host.info.n_out_samples, M = pl.host.info.n_mut, L = pl.host.info.n_loci;
  require(S * M == 0 || (d_depth && d_occ && d_cov), "table pointers are NULL");
  const double t0 = now_ms();
  if (!stats) {  // asynchronous, no table checksums: just the coverage gather, queued
    CUDA_OK(pcs::launch_finalize(st, d_depth, fo.d_row_locus.p, static_cast<uint32_t>(S), static_cast<uint32_t>(L),
                                 static_cast<uint32_t>(M), d_cov));
    pl.launches_since_read += S * M != 0 ? 1 : 0;
    return;
  }
  CUDA_OK(cudaMemsetAsync(pl.d_counters.p, 0, 4 * sizeof(unsigned long long), st));
  CUDA_OK(pcs::launch_finalize(st, d_depth, fo.d_row_locus.p, static_cast<uint32_t>(S), static_cast<uint32_t>(L),
                               static_cast<uint32_t>(M), d_cov));
  CUDA_OK(pcs::launch_sum_u32(st, d_depth, S * L, pl.d_counters.p + 1));
  CUDA_OK(pcs::launch_sum_u32(st, d_occ, S * M, pl.d_counters.p + 2));
  unsigned long long counters[4] = {0, 0, 0, 0};
  CUDA_OK(cudaMemcpyAsync(counters, pl.d_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  if (stats) {
    *stats = pcs_run_stats{};
    stats->total_ms = now_ms() - t0;
    stats->kernel_launches = (S * M != 0 ? 2 : 0) + (S * L != 0 ? 1 : 0);
    stats->sum_depth = counters[1];
    stats->sum_occurrences = counters[2];
    stats->d2h_bytes = sizeof(counters);
  }
}

// ------------------------------------------------------------ device -> caller copies
// Device memory to the caller's (pageable) buffers: the DMA lands in pinned staging chunk by chunk and host
// threads copy chunk k out while chunk k+1 is still on the link.
struct OutCopy {
  void* dst;
  const void* src;
  size_t bytes;
};

struct ChunkCopy {
  char* dst;
  const char* src;
  size_t bytes;
};

// wait for the chunks' events (cx.chunk_ev[k], recorded by the caller) and copy them out with host threads
void drain_chunks(pcs_ctx& cx, const std::vector<ChunkCopy>& chunks) {
  if (chunks.empty()) return;
  const unsigned nt = std::max(1u, std::min(16u, host_threads()));
  std::vector<cudaError_t> werr(nt, cudaSuccess);
  pcs::HostPool::get().run(nt, [&](size_t w) {  // thread w copies its share of every chunk, chunk after chunk
    cudaSetDevice(cx.device);
    for (size_t k = 0; k < chunks.size(); ++k) {
      const cudaError_t e = cudaEventSynchronize(cx.chunk_ev[k]);
      if (e != cudaSuccess) {
        werr[w] = e;
        return;
      }
      const size_t lo = (chunks[k].bytes * w / nt) & ~static_cast<size_t>(63);
      const size_t hi = w + 1 == nt ? chunks[k].bytes : (chunks[k].bytes * (w + 1) / nt) & ~static_cast<size_t>(63);
      if (hi > lo) std::memcpy(chunks[k].dst + lo, chunks[k].src + lo, hi - lo);
    }
  }, nt);
  for (cudaError_t e : werr) CUDA_OK(e);
}

// the same with one event per chunk, whatever device recorded it (pcs_simulate_multi: every device sends a slice)
void drain_chunks_ev(int device, const std::vector<ChunkCopy>& chunks, const std::vector<cudaEvent_t>& done) {
  if (chunks.empty()) return;
  const unsigned nt = std::max(1u, std::min(16u, host_threads()));
  std::vector<cudaError_t> werr(nt, cudaSuccess);
  pcs::HostPool::get().run(nt, [&](size_t w) {
    cudaSetDevice(device);
    for (size_t k = 0; k < chunks.size(); ++k) {
      const cudaError_t e = cudaEventSynchronize(done[k]);
      if (e != cudaSuccess) {
        werr[w] = e;
        return;
      }
      const size_t lo = (chunks[k].bytes * w / nt) & ~static_cast<size_t>(63);
      const size_t hi = w + 1 == nt ? chunks[k].bytes : (chunks[k].bytes * (w + 1) / nt) & ~static_cast<size_t>(63);
      if (hi > lo) std::memcpy(chunks[k].dst + lo, chunks[k].src + lo, hi - lo);
    }
  }, nt);
  for (cudaError_t e : werr) CUDA_OK(e);
}

uint64_t copy_out(pcs_ctx& cx, cudaStream_t st, const std::vector<OutCopy>& items) {
  size_t total = 0;
  for (const auto& it : items) total += (it.bytes + 255) & ~static_cast<size_t>(255);
  if (total == 0) return 0;
  for (const auto& it : items) advise_huge_pages(it.dst, it.bytes);
  char* stage = static_cast<char*>(cx.staging(total));
  const size_t cb = std::max<size_t>(4u << 20, ((total / 16) + 4095) & ~static_cast<size_t>(4095));
  std::vector<ChunkCopy> chunks;
  size_t at = 0;
  for (const auto& it : items) {
    for (size_t off = 0; off < it.bytes; off += cb) {
      const size_t nb = std::min(cb, it.bytes - off);
      CUDA_OK(cudaMemcpyAsync(stage + at + off, static_cast<const char*>(it.src) + off, nb, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaEventRecord(cx.chunk_event(chunks.size()), st));
      chunks.push_back({static_cast<char*>(it.dst) + off, stage + at + off, nb});
    }
    at += (it.bytes + 255) & ~static_cast<size_t>(255);
  }
  drain_chunks(cx, chunks);
  CUDA_OK(cudaStreamSynchronize(st));
  uint64_t bytes = 0;
  for (const auto& it : items) bytes += it.bytes;
  return bytes;
}

// ------------------------------------------------------------ one call, sample by sample
// pcs_simulate() and pcs_simulate_result(): the call is planned AND launched one output sample at a time, so the
// host plans sample s+1 (entries, multinomial template counts, tile order) while the GPU samples s -- of the
// planner's time only the first sample's share stays in front of the kernels.  A sample's slice of the plan goes
// up through a pinned slot with one DMA per array.  Same tile ids, template counts and Philox counters as the
// one-piece plan of pcs_plan_create(): the tables are identical, bit for bit.
struct CallTables {
  DevBuf<uint32_t> depth, occ, cov;
  DevBuf<unsigned long long> counters;
  DevBuf<uint32_t> insert_alias;
  size_t S = 0, M = 0, L = 0;
};

// after_part(sample, first row, end row, launches so far): called when the launches that write those rows of the
// sample's tables are queued
template <class AfterPart>
void run_pipelined(pcs_forest& fo, const pcs_seq_params& P, bool want_cov, bool split_last, CallTables& T, pcs_run_stats& rs,
                   AfterPart&& after_part) {
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  Lap lap;
  fo.sync_groups();
  const PlanSetup ps = plan_setup(fo.host, P);
  lap("  plan: tile geometry");
  const pcs::FlatForest& F = fo.host.flat;
  const size_t S = T.S = ps.samples.size(), M = T.M = F.n_mut, L = T.L = F.locus_pos.size();
  T.depth.alloc(S * L, st);
  T.occ.alloc(S * M, st);
  if (want_cov) T.cov.alloc(S * M, st);
  T.counters.alloc(4, st);
  if (S * L != 0) CUDA_OK(cudaMemsetAsync(T.depth.p, 0, T.depth.bytes(), st));
  if (S * M != 0) CUDA_OK(cudaMemsetAsync(T.occ.p, 0, T.occ.bytes(), st));
  CUDA_OK(cudaMemsetAsync(T.counters.p, 0, T.counters.bytes(), st));
  uint64_t h2d = 0, launches = 0, templates = 0;
  if (!ps.insert_alias.empty()) h2d += T.insert_alias.upload(ps.insert_alias, st);
  const pcs::DevForest DF = fo.dev();
  uint32_t id_base = 0;
  rs = pcs_run_stats{};
  for (uint32_t smp = 0; smp < S; ++smp) {
    HostPlan hp = make_host_plan(ps, smp, smp + 1, id_base);
    id_base += static_cast<uint32_t>(hp.info.n_tiles_total);
    templates += hp.info.n_templates;
    hp.model.insert_alias = T.insert_alias.p;
    // the slice goes up from one pinned slot, one DMA per array; the buffers are freed in stream order
    auto padded = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t b_tiles = hp.tiles.size() * sizeof(pcs::Tile), b_glob = hp.tiles_global.size() * sizeof(pcs::Tile),
                 b_ent = hp.entries.size() * sizeof(pcs::Entry), b_lo = hp.entry_lo.size() * sizeof(uint32_t);
    pcs_ctx::PlanSlot& slot = cx.plan_slot(smp % 4, padded(b_tiles) + padded(b_glob) + padded(b_ent) + padded(b_lo) + 256);
    DevBuf<pcs::Tile> d_tiles, d_glob;
    DevBuf<pcs::Entry> d_ent;
    DevBuf<uint32_t> d_lo;
    size_t at = 0;
    auto up = [&](auto& buf, const auto& vec, size_t bytes) {
      buf.alloc(vec.size(), st);
      if (bytes) {
        std::memcpy(slot.p + at, vec.data(), bytes);
        CUDA_OK(cudaMemcpyAsync(buf.p, slot.p + at, bytes, cudaMemcpyHostToDevice, st));
      }
      at += padded(bytes);
      h2d += bytes;
    };
    // The LAST sample's tables are the tail of the call: nothing is left to sample while they cross the link.  So
    // its tiles are launched in `parts` groups of whole chromosomes (rows of a chromosome are contiguous and a tile
    // never leaves its chromosome): the rows of a group travel while the next group is sampled, and the tail is
    // one group's tables instead of the sample's.  Same tiles, same counters: the tables do not change.
    std::vector<uint32_t> part_tile_end{static_cast<uint32_t>(hp.tiles.size())};  // staged tiles of parts 0..k
    std::vector<uint32_t> part_row_end{static_cast<uint32_t>(M)};
    // PCS_SPLIT_LAST=force: also on small jobs (tests); PCS_NO_SPLIT: never
    static const bool force_split = [] {
      const char* e = std::getenv("PCS_SPLIT_LAST");
      return e && std::string(e) == "force";
    }();
    if (split_last && smp + 1 == S && hp.tiles_global.empty() && F.n_chr >= 2 &&
        (force_split || (hp.tiles.size() >= 3000 && M >= (1u << 20)))) {
      constexpr uint32_t kParts = 3;
      std::vector<uint32_t> chr_part(F.n_chr, kParts - 1);
      part_row_end.assign(kParts, static_cast<uint32_t>(M));
      uint32_t c = 0;
      for (uint32_t k = 0; k + 1 < kParts; ++k) {  // chromosomes up to the one where the rows pass (k + 1) / kParts
        while (c < F.n_chr && F.locus_first_row[F.chr_locus_off[c + 1]] <= static_cast<uint64_t>(M) * (k + 1) / kParts) chr_part[c++] = k;
        part_row_end[k] = F.locus_first_row[F.chr_locus_off[c]];
      }
      std::vector<pcs::Tile> grouped;
      grouped.reserve(hp.tiles.size());
      part_tile_end.assign(kParts, 0);
      for (uint32_t k = 0; k < kParts; ++k) {  // heaviest first inside every group, as in the whole list
        for (const pcs::Tile& t : hp.tiles)
          if (chr_part[t.chr] == k) grouped.push_back(t);
        part_tile_end[k] = static_cast<uint32_t>(grouped.size());
      }
      hp.tiles.swap(grouped);
    }
    up(d_tiles, hp.tiles, b_tiles);
    up(d_glob, hp.tiles_global, b_glob);
    up(d_ent, hp.entries, b_ent);
    up(d_lo, hp.entry_lo, b_lo);
    CUDA_OK(cudaEventRecord(slot.done, st));
    slot.in_flight = true;
    CUDA_OK(cudaEventRecord(cx.time_event(2 * smp), st));
    uint32_t t0 = 0, r0 = 0;
    for (size_t k = 0; k < part_tile_end.size(); ++k) {
      CUDA_OK(pcs::launch_sample_tiles_staged(st, d_tiles.p + t0, part_tile_end[k] - t0, d_ent.p, d_lo.p, DF, hp.model,
                                              hp.dims, T.depth.p, T.occ.p, T.counters.p));
      launches += part_tile_end[k] > t0 ? 1 : 0;
      if (k + 1 == part_tile_end.size()) {
        CUDA_OK(pcs::launch_sample_tiles_global(st, d_glob.p, static_cast<uint32_t>(hp.tiles_global.size()), d_ent.p, d_lo.p,
                                                DF, hp.model, T.depth.p, T.occ.p, T.counters.p));
        launches += hp.tiles_global.empty() ? 0 : 1;
        CUDA_OK(cudaEventRecord(cx.time_event(2 * smp + 1), st));
      }
      after_part(smp, r0, part_row_end[k], launches);
      t0 = part_tile_end[k];
      r0 = part_row_end[k];
    }
    rs.n_templates += hp.info.n_templates;
    if (smp == 0) {
      rs.h2d_bytes = 0;  // filled below
      if (lap.on)
        std::fprintf(stderr, "[pcs host]    sample 0: %zu staged tiles, %zu global; staged smem %zu B\n", hp.tiles.size(),
                     hp.tiles_global.size(), pcs::staged_smem_bytes(hp.dims, hp.model.sequencer != PCS_SEQ_ERRORLESS));
    }
  }
  lap("  plan + launch, all samples");
  rs.kernel_launches = launches;
  rs.h2d_bytes = h2d;
  (void)templates;
}

// kernel time of a pipelined call: the per-sample sampler launches, summed (events of run_pipelined)
double pipelined_kernel_ms(pcs_ctx& cx, size_t S) {
  double total = 0;
  for (size_t s = 0; s < S; ++s) {
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, cx.time_ev[2 * s], cx.time_ev[2 * s + 1]));
    total += ms;
  }
  return total;
}

// pcs_simulate(): full tables back to the caller.  The tables of sample s cross the link on the copy stream while
// sample s+1 is being sampled.
void simulate_tables(pcs_forest& fo, const pcs_seq_params& P, uint32_t* occ, uint32_t* cov, pcs_run_stats* stats) {
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  cudaStream_t cs = cx.copier();
  const double t0 = now_ms();
  CallTables T;
  pcs_run_stats rs{};
  std::vector<ChunkCopy> chunks;
  char* stage = nullptr;
  const size_t M = fo.host.flat.n_mut, L = fo.host.flat.locus_pos.size();
  const size_t row_bytes = M * sizeof(uint32_t);
  {
    size_t n_out = (P.normal_only ? 0 : fo.host.n_groups) + ((P.normal_only || P.with_normal_sample) ? 1 : 0);
    advise_huge_pages(occ, n_out * row_bytes);
    advise_huge_pages(cov, n_out * row_bytes);
  }
  uint64_t fin_launches = 0;
  size_t n_parts = 0;
  run_pipelined(fo, P, true, std::getenv("PCS_NO_SPLIT") == nullptr, T, rs, [&](uint32_t smp, uint32_t r0, uint32_t r1, uint64_t) {
    if (M == 0 || r1 <= r0) return;
    require(occ && cov, "occurrences/coverage output pointers are NULL");
    if (!stage) stage = static_cast<char*>(cx.staging(2 * T.S * row_bytes));
    const size_t n_rows = r1 - r0, part_bytes = n_rows * sizeof(uint32_t);
    CUDA_OK(pcs::launch_finalize(st, T.depth.p + smp * L, fo.d_row_locus.p + r0, 1u, static_cast<uint32_t>(L),
                                 static_cast<uint32_t>(n_rows), T.cov.p + smp * M + r0));
    ++fin_launches;
    cudaEvent_t done = cx.chunk_event(2 * T.S + 8 + n_parts++);
    CUDA_OK(cudaEventRecord(done, st));
    CUDA_OK(cudaStreamWaitEvent(cs, done, 0));
    const char* dev_tbl[2] = {reinterpret_cast<const char*>(T.occ.p + smp * M + r0), reinterpret_cast<const char*>(T.cov.p + smp * M + r0)};
    char* host_tbl[2] = {reinterpret_cast<char*>(occ + smp * M + r0), reinterpret_cast<char*>(cov + smp * M + r0)};
    for (int t = 0; t < 2; ++t) {
      char* sp = stage + (2 * smp + t) * row_bytes + static_cast<size_t>(r0) * sizeof(uint32_t);
      CUDA_OK(cudaMemcpyAsync(sp, dev_tbl[t], part_bytes, cudaMemcpyDeviceToHost, cs));
      CUDA_OK(cudaEventRecord(cx.chunk_event(chunks.size()), cs));
      chunks.push_back({host_tbl[t], sp, part_bytes});
    }
  });
  if (stats) {
    CUDA_OK(pcs::launch_sum_u32(st, T.depth.p, T.S * L, T.counters.p + 1));
    CUDA_OK(pcs::launch_sum_u32(st, T.occ.p, T.S * M, T.counters.p + 2));
    fin_launches += (T.S * M != 0 ? 1 : 0) + (T.S * L != 0 ? 1 : 0);
  }
  unsigned long long* counters = cx.counters_home();
  CUDA_OK(cudaMemcpyAsync(counters, T.counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  Lap lap;
  drain_chunks(cx, chunks);
  lap("  tables copied out (waits for the GPU)");
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaStreamSynchronize(cs));
  lap("  streams drained");
  if (std::getenv("PCS_TIMING") && T.S > 0) {
    std::fprintf(stderr, "[pcs host]    %-28s %8.2f ms\n", "simulate (plan + kernels + D2H)", now_ms() - t0);
    float first = 0, all = 0;  // since the first sampler launch: when the first and the last sampler kernel ended
    cudaEventElapsedTime(&first, cx.time_ev[0], cx.time_ev[1]);
    cudaEventElapsedTime(&all, cx.time_ev[0], cx.time_ev[2 * T.S - 1]);
    std::fprintf(stderr, "[pcs host]      GPU: first sampler %.2f ms, first launch -> last sampler done %.2f ms, kernels summed %.2f ms\n",
                 first, all, pipelined_kernel_ms(cx, T.S));
  }
  if (stats) {
    rs.kernel_ms = pipelined_kernel_ms(cx, T.S);
    rs.total_ms = now_ms() - t0;
    rs.kernel_launches += fin_launches;
    rs.n_reads = counters[0];
    rs.sum_depth = counters[1];
    rs.sum_occurrences = counters[2];
    rs.d2h_bytes = 2 * T.S * row_bytes + 4 * sizeof(unsigned long long);
    *stats = rs;
  }
}

// rows some SEQUENCED cell inherits (include_non_sequenced_mutations): an instance whose haplotype interval holds
// a haplotype of a sequenced kind (tumour cells unless normal_only; the normal cells that are in the mix) that
// still holds the position, on a sequenced chromosome
std::vector<uint8_t> carried_rows(const HostForest& host, const pcs_seq_params* params) {
  const pcs::FlatForest& F = host.flat;
  const bool normal_only = params && params->normal_only;
  const bool preneo = params && params->preneoplastic_in_normal;
  const bool normals = !params || normal_only || params->with_normal_sample || params->purity < 1.0;
  std::vector<uint8_t> carried(F.n_mut, 0);
  host_tasks(F.n_chr, [&](size_t ci) {
    const uint32_t c = static_cast<uint32_t>(ci);
    // the reference hands only the chromosomes of `chr_ids` to the simulator (src/seq_simulation.cpp:570,577):
    // rows of the others never reach its data frame
    if (params && params->chr_mask && !params->chr_mask[c]) return;
    std::vector<uint32_t> seq_haps;  // sorted haplotype indices of the sequenced cells
    const auto& haps = F.chr_haps[c];
    for (uint32_t h = 0; h < haps.size(); ++h) {
      const bool on = haps[h].kind == pcs::HAP_TUMOUR ? !normal_only
                      : normals && (haps[h].kind == (preneo ? pcs::HAP_NORMAL_PRENEO : pcs::HAP_NORMAL_PLAIN));
      if (on) seq_haps.push_back(h);
    }
    for (uint32_t l = F.chr_locus_off[c]; l < F.chr_locus_off[c + 1]; ++l)
      for (uint32_t i = F.locus_inst_off[l]; i < F.locus_inst_off[l + 1]; ++i) {
        const pcs::Inst& in = F.inst[i];
        if (carried[in.row]) continue;
        // a haplotype of the interval that still HOLDS the position (a later CNA deletion may have taken it)
        const uint32_t pos = F.locus_pos[l];
        for (auto it = std::lower_bound(seq_haps.begin(), seq_haps.end(), in.lo);
             it != seq_haps.end() && *it - in.lo < in.span; ++it) {
          bool has = false;
          for (const auto& fr : F.fragsets[haps[*it].fragset]) has |= pos >= fr.b && pos <= fr.e;
          if (has) {
            carried[in.row] = 1;
            break;
          }
        }
      }
  });
  return carried;
}

// active rows of the tables (d_occ, d_depth) compacted on the device into a pcs_result
std::unique_ptr<pcs_result> assemble_result(pcs_forest& fo, const uint32_t* d_occ, const uint32_t* d_depth, size_t S,
                                            int include_non_sequenced, const pcs_seq_params* params, bool want_vaf) {
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  const pcs::FlatForest& F = fo.host.flat;
  const uint32_t M = F.n_mut, L = static_cast<uint32_t>(F.locus_pos.size());
  auto res = std::make_unique<pcs_result>();
  res->forest = &fo;
  res->n_samples = static_cast<uint32_t>(S);
  if (M == 0 || S == 0) return res;
  DevBuf<uint8_t> d_carried;
  if (include_non_sequenced) {
    fo.ensure_host_instances();
    const std::vector<uint8_t> carried = carried_rows(fo.host, params);
    d_carried.upload(carried, st);
  }
  DevBuf<uint32_t> d_blocks;
  d_blocks.alloc(static_cast<size_t>(pcs::active_blocks(M)) + 1, st);
  uint32_t* d_total = d_blocks.p + pcs::active_blocks(M);
  CUDA_OK(pcs::launch_active_count(st, d_occ, d_carried.p, static_cast<uint32_t>(S), M, d_blocks.p, d_total));
  unsigned long long* home = cx.counters_home();
  uint32_t* n_home = reinterpret_cast<uint32_t*>(home + 3);  // the trace counter's slot: unused on this path
  CUDA_OK(cudaMemcpyAsync(n_home, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  const uint32_t n = *n_home;
  res->n_rows = n;
  res->d_rows.alloc(n, st);
  res->d_occ.alloc(S * n, st);
  res->d_cov.alloc(S * n, st);
  if (want_vaf) res->d_vaf.alloc(S * n, st);
  CUDA_OK(pcs::launch_active_scatter(st, d_occ, d_depth, fo.d_row_locus.p, d_carried.p, static_cast<uint32_t>(S), M, L,
                                     d_blocks.p, n, res->d_rows.p, res->d_occ.p, res->d_cov.p, res->d_vaf.p));
  return res;
}

// reads of a set of tiles as binary records (host vectors)
struct Materialized {
  std::vector<pcs::SamHeader> hdr;
  std::vector<uint32_t> masks;
  std::vector<uint8_t> seq, qual;
};

void materialize_tiles(pcs_plan& pl, const std::vector<pcs::Tile>& tiles, uint64_t cap, Materialized& out) {
  pcs_forest& fo = *pl.forest;
  pcs_ctx& cx = *fo.ctx;
  cx.bind();
  cudaStream_t st = cx.stream;
  const uint32_t R = pl.host.info.read_size;
  for (const auto& t : tiles) require(fo.has_reference(t.chr), "the reference sequence of a sequenced chromosome is not loaded");
  const pcs::SeqData D = fo.seq_data();
  DevBuf<pcs::Tile> d_tiles;
  DevBuf<pcs::SamHeader> d_hdr;
  DevBuf<uint32_t> d_masks;
  DevBuf<uint8_t> d_seq, d_qual;
  d_tiles.upload(tiles, st);
  d_hdr.alloc(cap, st);
  d_masks.alloc(cap * PCS_ERRMASK_WORDS, st);
  d_seq.alloc(cap * R, st);
  d_qual.alloc(cap * R, st);
  CUDA_OK(cudaMemsetAsync(pl.d_counters.p, 0, 4 * sizeof(unsigned long long), st));
  CUDA_OK(pcs::launch_materialize_tiles(st, d_tiles.p, static_cast<uint32_t>(tiles.size()), pl.d_entries.p, pl.d_entry_lo.p, fo.dev(),
                                        pl.host.model, D, d_hdr.p, d_masks.p, d_seq.p, d_qual.p, cap, pl.d_counters.p + 3));
  unsigned long long counters[4];
  CUDA_OK(cudaMemcpyAsync(counters, pl.d_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  if (counters[3] > cap) throw std::domain_error("materialize capacity too small");
  const size_t n = counters[3];
  out.hdr.resize(n);
  out.masks.resize(n * PCS_ERRMASK_WORDS);
  out.seq.resize(n * R);
  out.qual.resize(n * R);
  if (n) {
    CUDA_OK(cudaMemcpyAsync(out.hdr.data(), d_hdr.p, n * sizeof(pcs::SamHeader), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(out.masks.data(), d_masks.p, out.masks.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(out.seq.data(), d_seq.p, out.seq.size(), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(out.qual.data(), d_qual.p, out.qual.size(), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  // a CIGAR the record cannot hold would no longer describe SEQ: refuse, never truncate
  for (const auto& h : out.hdr)
    if (h.flags & 4u)
      throw std::domain_error("a read carries more indels than a CIGAR of " + std::to_string(pcs::kMaxCigar) +
                              " operations can describe (tile " + std::to_string(h.tile_id) + ", read " +
                              std::to_string(h.read_id) + ")");
}

void placement_of(const pcs::FlatForest& F, const pcs::SamHeader& h, pcs_read_placement& r) {
  const uint32_t chr = h.chr_sample & 0xffffu;
  const pcs::HapRec& hr = F.chr_haps[chr][h.hap];
  r.cell = hr.cell;
  r.start = h.start;
  r.chr = static_cast<uint16_t>(chr);
  r.allele = hr.allele;
  r.sample = static_cast<uint16_t>(h.chr_sample >> 16);
  r.flags = hr.kind == pcs::HAP_TUMOUR ? PCS_PLACE_TUMOUR
            : hr.kind == pcs::HAP_NORMAL_PLAIN ? PCS_PLACE_NORMAL_PLAIN : PCS_PLACE_NORMAL_PRENEO;
}

void append_uint(std::string& s, uint64_t v) {
  char buf[24];
  int n = 0;
  do { buf[n++] = static_cast<char>('0' + v % 10); v /= 10; } while (v);
  while (n) s.push_back(buf[--n]);
}

// SAM text of one record.  Template names: <prefix><tile>_<template index> (unique per file, shared by mates).
void append_sam_line(std::string& s, const pcs::SamHeader& h, const uint8_t* seq, const uint8_t* qual,
                     const std::string& prefix, const std::string& chr, const std::string& sample) {
  const bool paired = (h.flags & 1u) != 0, second = (h.flags & 2u) != 0;
  s += prefix;
  append_uint(s, h.tile_id);
  s.push_back('_');
  append_uint(s, paired ? h.read_id >> 1 : h.read_id);
  s.push_back('\t');
  append_uint(s, paired ? (second ? 147u : 99u) : 0u);
  s.push_back('\t');
  s += chr;
  s.push_back('\t');
  append_uint(s, h.start);
  s += "\t60\t";
  if (h.n_cigar == 0) s.push_back('*');
  for (uint32_t i = 0; i < h.n_cigar; ++i) {
    append_uint(s, h.cigar[i] >> 4);
    s.push_back("MID"[h.cigar[i] & 3u]);
  }
  s.push_back('\t');
  if (paired) {
    s += "=\t";
    append_uint(s, h.mate_start);
    s.push_back('\t');
    if (h.tlen < 0) s.push_back('-');
    append_uint(s, static_cast<uint64_t>(h.tlen < 0 ? -static_cast<int64_t>(h.tlen) : h.tlen));
  } else {
    s += "*\t0\t0";
  }
  s.push_back('\t');
  s.append(reinterpret_cast<const char*>(seq), h.len);
  s.push_back('\t');
  s.append(reinterpret_cast<const char*>(qual), h.len);
  s += "\tRG:Z:";
  s += sample;
  s.push_back('\n');
}

// peer access between two devices of this process, both directions (idempotent)
void enable_peer_both_ways(int a, int b) {
  if (a == b) return;
  static std::mutex mu;
  static std::vector<std::pair<int, int>> done;  // pairs already set up by this process
  std::lock_guard<std::mutex> lock(mu);
  if (std::find(done.begin(), done.end(), std::make_pair(std::min(a, b), std::max(a, b))) != done.end()) return;
  const int pair[2][2] = {{a, b}, {b, a}};
  for (const auto& pr : pair) {
    CUDA_OK(cudaSetDevice(pr[0]));
    int can = 0;
    CUDA_OK(cudaDeviceCanAccessPeer(&can, pr[0], pr[1]));
    if (!can) throw CudaError("the devices cannot access each other's memory");
    const cudaError_t e = cudaDeviceEnablePeerAccess(pr[1], 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else CUDA_OK(e);
    // the library's buffers come from the stream-ordered pool, which cudaDeviceEnablePeerAccess does not cover:
    // without this a device-to-device copy of pool memory is staged through the host
    cudaMemPool_t pool;
    CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, pr[0]));
    cudaMemAccessDesc desc{};
    desc.location.type = cudaMemLocationTypeDevice;
    desc.location.id = pr[1];
    desc.flags = cudaMemAccessFlagsProtReadWrite;
    CUDA_OK(cudaMemPoolSetAccess(pool, &desc, 1));
  }
  done.emplace_back(std::min(a, b), std::max(a, b));
}

template <class Fn>
int guarded(Fn&& fn) {
  try {
    fn();
    return PCS_OK;
  } catch (const CudaError& e) {
    g_err = e.what();
    return PCS_ERR_CUDA;
  } catch (const std::bad_alloc&) {
    g_err = "out of host memory";
    return PCS_ERR_NOMEM;
  } catch (const std::domain_error& e) {
    g_err = e.what();
    return PCS_ERR_INVALID;
  } catch (const std::exception& e) {
    g_err = e.what();
    return PCS_ERR_INTERNAL;
  }
}

}  // namespace

extern "C" {

int pcs_abi_version(void) { return PCS_ABI_VERSION; }

const char* pcs_last_error(void) { return g_err.c_str(); }

int pcs_create(pcs_ctx** out, int device_id, void* stream) {
  return guarded([&] {
    require(out != nullptr, "ctx output pointer is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw CudaError(std::string("no CUDA device available (libpcs_seq has no CPU path): ") +
                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    require(device_id >= 0 && device_id < n, "device_id out of range");
    auto cx = std::make_unique<pcs_ctx>();
    cx->device = device_id;
    cx->bind();
    cudaDeviceProp prop{};
    CUDA_OK(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10)
      throw CudaError(std::string("libpcs_seq is built for sm_100a only; device is ") + prop.name + " (sm_" +
                      std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    if (stream) {
      cx->stream = static_cast<cudaStream_t>(stream);
    } else {
      CUDA_OK(cudaStreamCreateWithFlags(&cx->stream, cudaStreamNonBlocking));
      cx->own_stream = true;
    }
    for (auto& ev : cx->ev) CUDA_OK(cudaEventCreate(&ev));
    cudaMemPool_t pool;
    CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, device_id));
    uint64_t keep = ~0ull;  // do not hand pooled memory back to the driver at every synchronisation
    CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    *out = cx.release();
  });
}

extern "C++" {
static void free_ctx(pcs_ctx* cx) {
    cudaSetDevice(cx->device);
    for (auto& ev : cx->ev)
      if (ev) cudaEventDestroy(ev);
    if (cx->own_stream && cx->stream) cudaStreamDestroy(cx->stream);
    if (cx->pinned) cudaFreeHost(cx->pinned);
    if (cx->pinned_counters) cudaFreeHost(cx->pinned_counters);
    if (cx->peer_tables) cudaFree(cx->peer_tables);
    for (auto& b : cx->pool) cudaFreeHost(b.p);
    for (auto& ev : cx->chunk_ev) cudaEventDestroy(ev);
    for (auto& ev : cx->time_ev) cudaEventDestroy(ev);
    for (auto& sl : cx->plan_slots) {
      if (sl.p) cudaFreeHost(sl.p);
      if (sl.done) cudaEventDestroy(sl.done);
    }
    if (cx->copy_stream) cudaStreamDestroy(cx->copy_stream);
    delete cx;
}

void release_ctx(pcs_ctx* cx) {
  if (--cx->users == 0 && cx->destroyed.load()) free_ctx(cx);
}
}  // extern "C++"

int pcs_destroy(pcs_ctx* cx) {
  return guarded([&] {
    if (!cx) return;
    cx->destroyed.store(true);
    if (cx->users.load() == 0) free_ctx(cx);  // else: the last forest to be freed does it
  });
}

int pcs_device_name(pcs_ctx* cx, char* buf, size_t len) {
  return guarded([&] {
    require(cx && buf && len > 0, "bad arguments");
    cudaDeviceProp prop{};
    CUDA_OK(cudaGetDeviceProperties(&prop, cx->device));
    std::strncpy(buf, prop.name, len - 1);
    buf[len - 1] = 0;
  });
}

int pcs_forest_upload(pcs_ctx* cx, const pcs_forest_desc* desc, pcs_forest** out) {
  return guarded([&] {
    require(cx && desc && out, "bad arguments");
    require(!cx->destroyed.load(), "the context has been destroyed");
    auto fo = std::make_unique<pcs_forest>();
    fo->ctx = cx;
    fo->user.hold(cx);
    Lap lap;
    fo->host.borrow(cx, pcs::flat_store_bytes(*desc));
    lap("pinned block");
    pcs_forest* early = fo.get();
    pcs::flatten_forest(*desc, fo->host.flat, host_threads(), [early] { early->upload_loci_early(); }, device_instances());
    lap("flatten_forest");
    fo->upload_flat();
    lap("upload flat arrays");
    fo->set_groups(fo->host.flat.leaf_sample.data(), fo->host.flat.n_samples);
    lap("groups + upload");
    *out = fo.release();
  });
}

int pcs_forest_upload_genomes(pcs_ctx* cx, const pcs_cell_genomes_desc* desc, pcs_forest** out) {
  return guarded([&] {
    require(cx && desc && out, "bad arguments");
    require(!cx->destroyed.load(), "the context has been destroyed");
    auto fo = std::make_unique<pcs_forest>();
    fo->ctx = cx;
    fo->user.hold(cx);
    Lap lap;
    fo->host.borrow(cx, pcs::flat_store_bytes(*desc));
    pcs_forest* early = fo.get();
    pcs::flatten_cell_genomes(*desc, fo->host.flat, host_threads(), [early] { early->upload_loci_early(); }, device_instances());
    lap("flatten_cell_genomes");
    fo->upload_flat();
    fo->set_groups(fo->host.flat.leaf_sample.data(), fo->host.flat.n_samples);
    lap("upload");
    *out = fo.release();
  });
}

int pcs_forest_free(pcs_forest* fo) {
  return guarded([&] {
    if (!fo) return;
    pcs_ctx* cx = fo->ctx;
    cudaSetDevice(cx->device);
    cudaStreamSynchronize(cx->stream);  // an upload may still be reading the forest's pinned block
    delete fo;
  });
}

int pcs_forest_set_groups(pcs_forest* fo, const uint32_t* leaf_group, uint32_t n_groups) {
  return guarded([&] {
    require(fo != nullptr, "forest is NULL");
    if (leaf_group)
      fo->set_groups(leaf_group, n_groups);
    else
      fo->set_groups(fo->host.flat.leaf_sample.data(), fo->host.flat.n_samples);
  });
}

int pcs_forest_info(const pcs_forest* fo, uint64_t out[6]) {
  return guarded([&] {
    require(fo && out, "bad arguments");
    uint64_t haps = 0;
    for (const auto& v : fo->host.flat.chr_haps) haps += v.size();
    out[0] = fo->host.flat.locus_pos.size();
    out[1] = fo->host.flat.n_inst;
    out[2] = haps;
    out[3] = fo->host.flat.fragsets.size();
    out[4] = fo->host.flat.pieces.size();
    out[5] = fo->d_locus_pos.bytes() + fo->d_chr_locus_off.bytes() + fo->d_locus_inst_off.bytes() +
             fo->d_row_locus.bytes() + fo->d_inst.bytes() + fo->d_hap_list.bytes();
  });
}

int pcs_forest_instances(pcs_forest* fo, uint32_t* inst, uint32_t* locus_inst_off) {
  return guarded([&] {
    require(fo && inst && locus_inst_off, "bad arguments");
    fo->ctx->bind();
    cudaStream_t st = fo->ctx->stream;
    if (fo->d_inst.n) CUDA_OK(cudaMemcpyAsync(inst, fo->d_inst.p, fo->d_inst.bytes(), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(locus_inst_off, fo->d_locus_inst_off.p, fo->d_locus_inst_off.bytes(), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  });
}

int pcs_plan_create(pcs_forest* fo, const pcs_seq_params* params, pcs_plan** out) {
  return guarded([&] {
    require(fo && params && out, "bad arguments");
    validate(*params);
    auto pl = std::make_unique<pcs_plan>();
    pl->forest = fo;
    Lap lap;
    fo->sync_groups();
    pl->host = make_host_plan(fo->host, *params);
    lap("make_host_plan");
    upload_plan(*pl);
    lap("upload plan + alloc tables");
    if (lap.on)
      std::fprintf(stderr, "[pcs host]    tiles: %zu staged, %zu global; staged smem %zu B\n", pl->host.tiles.size(),
                   pl->host.tiles_global.size(), pcs::staged_smem_bytes(pl->host.dims, pl->host.model.sequencer != PCS_SEQ_ERRORLESS));
    *out = pl.release();
  });
}

int pcs_plan_info_get(const pcs_plan* pl, pcs_plan_info* info) {
  return guarded([&] {
    require(pl && info, "bad arguments");
    *info = pl->host.info;
  });
}

int pcs_plan_free(pcs_plan* pl) {
  return guarded([&] {
    if (!pl) return;
    cudaSetDevice(pl->forest->ctx->device);
    delete pl;
  });
}

int pcs_plan_run(pcs_plan* pl, int flags, uint32_t* occ, uint32_t* cov, pcs_run_stats* stats) {
  return guarded([&] {
    require(pl != nullptr, "plan is NULL");
    run_plan(*pl, flags, occ, cov, stats);
  });
}

int pcs_plan_accumulate(pcs_plan* pl, uint32_t* depth, uint32_t* occurrences, pcs_run_stats* stats) {
  return guarded([&] {
    require(pl != nullptr, "plan is NULL");
    accumulate_plan(*pl, depth, occurrences, stats);
  });
}

int pcs_plan_finalize(pcs_plan* pl, const uint32_t* depth, const uint32_t* occurrences, uint32_t* coverage,
                      pcs_run_stats* stats) {
  return guarded([&] {
    require(pl != nullptr, "plan is NULL");
    finalize_tables(*pl, depth, occurrences, coverage, stats);
  });
}

int pcs_plan_finalize_stream(pcs_plan* pl, const uint32_t* depth, uint32_t* coverage, void* stream) {
  return guarded([&] {
    require(pl != nullptr, "plan is NULL");
    pcs_forest& fo = *pl->forest;
    fo.ctx->bind();
    const size_t S = pl->host.info.n_out_samples, M = pl->host.info.n_mut, L = pl->host.info.n_loci;
    require(S * M == 0 || (depth && coverage), "table pointers are NULL");
    CUDA_OK(pcs::launch_finalize(static_cast<cudaStream_t>(stream), depth, fo.d_row_locus.p, static_cast<uint32_t>(S),
                                 static_cast<uint32_t>(L), static_cast<uint32_t>(M), coverage));
    pl->launches_since_read += S * M != 0 ? 1 : 0;
  });
}

int pcs_plan_counters(pcs_plan* pl, pcs_run_stats* stats) {
  return guarded([&] {
    require(pl && stats, "bad arguments");
    pcs_ctx& cx = *pl->forest->ctx;
    cx.bind();
    cudaStream_t st = cx.stream;
    unsigned long long* counters = cx.counters_home();
    CUDA_OK(cudaMemcpyAsync(counters, pl->d_counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemsetAsync(pl->d_counters.p, 0, 4 * sizeof(unsigned long long), st));
    CUDA_OK(cudaStreamSynchronize(st));
    *stats = pcs_run_stats{};
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, cx.ev[1], cx.ev[2]));  // the last sampler launch
    stats->kernel_ms = ms;
    stats->kernel_launches = pl->launches_since_read;
    stats->n_templates = pl->host.info.n_templates;
    stats->n_reads = counters[0];
    stats->sum_depth = counters[1];
    stats->sum_occurrences = counters[2];
    stats->d2h_bytes = 4 * sizeof(unsigned long long);
    pl->launches_since_read = 0;
  });
}

int pcs_memset_u32_stream(pcs_ctx* cx, uint32_t* dev_ptr, size_t count, void* stream) {
  return guarded([&] {
    require(cx && (dev_ptr || count == 0), "bad arguments");
    cx->bind();
    if (count) CUDA_OK(cudaMemsetAsync(dev_ptr, 0, count * sizeof(uint32_t), static_cast<cudaStream_t>(stream)));
  });
}

int pcs_shared_alloc(pcs_ctx* cx, size_t bytes, void** dev_ptr, unsigned char ipc_handle[64]) {
  return guarded([&] {
    require(cx && dev_ptr && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cx->bind();
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, bytes));  // plain cudaMalloc: pool memory cannot be exported this way
    if (ipc_handle) {
      cudaIpcMemHandle_t h;
      cudaError_t e = cudaIpcGetMemHandle(&h, p);
      if (e != cudaSuccess) {
        cudaFree(p);
        throw CudaError(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
      }
      std::memcpy(ipc_handle, &h, 64);
    }
    *dev_ptr = p;
  });
}

int pcs_shared_free(pcs_ctx* cx, void* dev_ptr) {
  return guarded([&] {
    require(cx != nullptr, "ctx is NULL");
    cx->bind();
    if (dev_ptr) CUDA_OK(cudaFree(dev_ptr));
  });
}

int pcs_shared_open(pcs_ctx* cx, const unsigned char ipc_handle[64], void** dev_ptr) {
  return guarded([&] {
    require(cx && ipc_handle && dev_ptr, "bad arguments");
    cx->bind();
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, 64);
    CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  });
}

int pcs_shared_close(pcs_ctx* cx, void* dev_ptr) {
  return guarded([&] {
    require(cx != nullptr, "ctx is NULL");
    cx->bind();
    if (dev_ptr) CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  });
}

int pcs_enable_peer(pcs_ctx* cx, int peer_device) {
  return guarded([&] {
    require(cx != nullptr, "ctx is NULL");
    cx->bind();
    if (peer_device == cx->device) return;
    int can = 0;
    CUDA_OK(cudaDeviceCanAccessPeer(&can, cx->device, peer_device));
    if (!can) throw CudaError("the devices cannot access each other's memory");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
      return;
    }
    CUDA_OK(e);
  });
}

int pcs_memset_u32(pcs_ctx* cx, uint32_t* dev_ptr, size_t count) {
  return guarded([&] {
    require(cx && (dev_ptr || count == 0), "bad arguments");
    cx->bind();
    if (count) CUDA_OK(cudaMemsetAsync(dev_ptr, 0, count * sizeof(uint32_t), cx->stream));
  });
}

int pcs_forest_replicate(pcs_forest* src, pcs_ctx* cx, pcs_forest** out) {
  return guarded([&] {
    require(src && cx && out, "bad arguments");
    require(!cx->destroyed.load(), "the context has been destroyed");
    Lap lap;
    auto fo = std::make_unique<pcs_forest>(*src, cx);  // shares the flattened host view: no second flatten
    fo->user.hold(cx);
    pcs_ctx& sx = *src->ctx;
    if (sx.device == cx->device) {
      fo->upload_flat();
      fo->upload_groups();
    } else {
      // device to device over NVLink: the tables are already in the source's HBM
      src->sync_groups();
      sx.bind();
      CUDA_OK(cudaStreamSynchronize(sx.stream));  // the source's own upload may still be in flight
      enable_peer_both_ways(sx.device, cx->device);
      cx->bind();
      cudaStream_t st = cx->stream;
      auto clone = [&](auto& dst, const auto& from) {
        dst.alloc(from.n, st);
        if (from.n) CUDA_OK(cudaMemcpyPeerAsync(dst.p, cx->device, from.p, sx.device, from.bytes(), st));
      };
      clone(fo->d_chr_locus_off, src->d_chr_locus_off);
      clone(fo->d_locus_pos, src->d_locus_pos);
      clone(fo->d_locus_inst_off, src->d_locus_inst_off);
      clone(fo->d_row_locus, src->d_row_locus);
      clone(fo->d_inst, src->d_inst);
      clone(fo->d_hap_list, src->d_hap_list);
      fo->uploaded_groups = src->uploaded_groups;
      CUDA_OK(cudaStreamSynchronize(st));
    }
    lap("replicate forest");
    *out = fo.release();
  });
}

// One process, several GPUs (the single-threaded R session).  The job is planned ONCE, sample by sample, for all
// shards (make_host_plans); the calling thread queues every device's slice and kernels on that device's stream --
// launches are asynchronous, no worker threads -- while it plans the next sample.  Every device's sampler adds
// straight into the tables on forests[0]'s device (peer memory over NVLink); when the last device has finished a
// sample, that sample's coverage gather and its copy to the host run on device 0's copy stream, behind the other
// samples' kernels.
int pcs_simulate_multi(pcs_forest* const* forests, uint32_t n, const pcs_seq_params* params, uint32_t* occ,
                       uint32_t* cov, pcs_run_stats* stats) {
  return guarded([&] {
    require(forests && n >= 1 && n <= 64 && params, "bad arguments");
    for (uint32_t i = 0; i < n; ++i) require(forests[i] != nullptr, "forest is NULL");
    for (uint32_t i = 1; i < n; ++i)
      require(forests[i]->host_ptr == forests[0]->host_ptr, "the forests must be replicas of forests[0] (pcs_forest_replicate)");
    validate(*params);
    const double t0 = now_ms();
    Lap lap;
    pcs_forest& owner = *forests[0];
    pcs_ctx& cx0 = *owner.ctx;
    pcs_seq_params P = *params;
    P.shard_rank = 0;
    P.shard_count = n;
    for (uint32_t i = 0; i < n; ++i) forests[i]->sync_groups();
    const PlanSetup ps = plan_setup(owner.host, P);
    lap("  plan: tile geometry");
    const pcs::FlatForest& F = owner.host.flat;
    const size_t S = ps.samples.size(), M = F.n_mut, L = F.locus_pos.size();
    require(S * M == 0 || (occ && cov), "occurrences/coverage output pointers are NULL");
    advise_huge_pages(occ, S * M * sizeof(uint32_t));
    advise_huge_pages(cov, S * M * sizeof(uint32_t));
    for (uint32_t i = 1; i < n; ++i) enable_peer_both_ways(cx0.device, forests[i]->ctx->device);
    // the tables live on the first device.  Plain cudaMalloc, kept by the context: peers cannot reach stream-ordered
    // pool memory through cudaDeviceEnablePeerAccess
    cx0.bind();
    const size_t words = S * L + 2 * S * M;
    uint32_t* tb = static_cast<uint32_t*>(cx0.peer_table(words * sizeof(uint32_t)));
    uint32_t* t_depth = tb;
    uint32_t* t_occ = tb + S * L;
    uint32_t* t_cov = t_occ + S * M;
    cudaStream_t st0 = cx0.stream, cs0 = cx0.copier();
    if (words) CUDA_OK(cudaMemsetAsync(tb, 0, (S * L + S * M) * sizeof(uint32_t), st0));
    // events live on the device whose stream records them (any stream may wait for them): in every context's pool,
    // [0, 2S) mark D2H chunks (context 0), 2S = tables zeroed (context 0), 2S + 1 + s = this device is through with sample s
    cudaEvent_t zeroed = cx0.chunk_event(2 * S);
    CUDA_OK(cudaEventRecord(zeroed, st0));
    std::vector<DevBuf<unsigned long long>> d_counters(n);
    std::vector<DevBuf<uint32_t>> d_alias(n);
    uint64_t h2d = 0, launches = 0;
    for (uint32_t i = 0; i < n; ++i) {
      pcs_ctx& cx = *forests[i]->ctx;
      cx.bind();
      d_counters[i].alloc(4, cx.stream);
      CUDA_OK(cudaMemsetAsync(d_counters[i].p, 0, 4 * sizeof(unsigned long long), cx.stream));
      if (!ps.insert_alias.empty()) h2d += d_alias[i].upload(ps.insert_alias, cx.stream);
      if (i) CUDA_OK(cudaStreamWaitEvent(cx.stream, zeroed, 0));
    }
    std::vector<ChunkCopy> chunks;
    std::vector<cudaEvent_t> chunk_done;
    const size_t row_bytes = M * sizeof(uint32_t);
    char* stage = row_bytes ? static_cast<char*>(cx0.staging(2 * S * row_bytes)) : nullptr;
    // Row slices of the tables, one per device whose link carries a share.  Default: ONE link, device 0's --
    // measured on the 8-GPU box (profiles/r02_v5_multi_links.md): with every device sending a slice the call is
    // 2-3 ms SLOWER: whatever bounds the copy-out there, it is not one device's PCIe link.  PCS_MULTI_LINKS=k: k
    // links; =force: n links whatever the size of the tables (tests).
    const uint32_t n_links = [&] {
      const char* e = std::getenv("PCS_MULTI_LINKS");
      if (e && std::string(e) == "force") return M >= 128 ? n : 1u;
      const long v = e ? std::atol(e) : 1;
      return M < (1u << 16) ? 1u : std::min<uint32_t>(n, static_cast<uint32_t>(std::max(1l, v)));
    }();
    auto slice_lo = [&](uint32_t i) { return i >= n_links ? M : (M * static_cast<size_t>(i) / n_links) & ~static_cast<size_t>(63); };
    size_t slice_cap = 0;
    for (uint32_t i = 1; i < n_links; ++i) slice_cap = std::max(slice_cap, slice_lo(i + 1) - slice_lo(i));
    std::vector<uint32_t*> scratch(n, nullptr);
    for (uint32_t i = 1; i < n_links; ++i) {
      forests[i]->ctx->bind();
      scratch[i] = static_cast<uint32_t*>(forests[i]->ctx->peer_table(2 * S * slice_cap * sizeof(uint32_t)));
    }
    cx0.bind();
    uint32_t id_base = 0;
    uint64_t templates = 0;
    auto padded = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    for (uint32_t smp = 0; smp < S; ++smp) {
      std::vector<HostPlan> plans = make_host_plans(ps, smp, smp + 1, id_base, true);
      id_base += static_cast<uint32_t>(plans[0].info.n_tiles_total);
      for (uint32_t i = 0; i < n; ++i) {
        HostPlan& hp = plans[i];
        pcs_forest& fo = *forests[i];
        pcs_ctx& cx = *fo.ctx;
        cx.bind();
        cudaStream_t st = cx.stream;
        templates += hp.info.n_templates;
        hp.model.insert_alias = d_alias[i].p;
        const size_t b_tiles = hp.tiles.size() * sizeof(pcs::Tile), b_glob = hp.tiles_global.size() * sizeof(pcs::Tile),
                     b_ent = hp.entries.size() * sizeof(pcs::Entry), b_lo = hp.entry_lo.size() * sizeof(uint32_t);
        pcs_ctx::PlanSlot& slot = cx.plan_slot(smp % 4, padded(b_tiles) + padded(b_glob) + padded(b_ent) + padded(b_lo) + 256);
        DevBuf<pcs::Tile> d_tiles, d_glob;
        DevBuf<pcs::Entry> d_ent;
        DevBuf<uint32_t> d_lo;
        size_t at = 0;
        auto up = [&](auto& buf, const auto& vec, size_t bytes) {
          buf.alloc(vec.size(), st);
          if (bytes) {
            std::memcpy(slot.p + at, vec.data(), bytes);
            CUDA_OK(cudaMemcpyAsync(buf.p, slot.p + at, bytes, cudaMemcpyHostToDevice, st));
          }
          at += padded(bytes);
          h2d += bytes;
        };
        up(d_tiles, hp.tiles, b_tiles);
        up(d_glob, hp.tiles_global, b_glob);
        up(d_ent, hp.entries, b_ent);
        up(d_lo, hp.entry_lo, b_lo);
        CUDA_OK(cudaEventRecord(slot.done, st));
        slot.in_flight = true;
        const pcs::DevForest DF = fo.dev();
        CUDA_OK(cudaEventRecord(cx.time_event(2 * smp), st));
        CUDA_OK(pcs::launch_sample_tiles_staged(st, d_tiles.p, static_cast<uint32_t>(hp.tiles.size()), d_ent.p, d_lo.p, DF,
                                                hp.model, hp.dims, t_depth, t_occ, d_counters[i].p));
        CUDA_OK(pcs::launch_sample_tiles_global(st, d_glob.p, static_cast<uint32_t>(hp.tiles_global.size()), d_ent.p,
                                                d_lo.p, DF, hp.model, t_depth, t_occ, d_counters[i].p));
        CUDA_OK(cudaEventRecord(cx.time_event(2 * smp + 1), st));
        launches += (hp.tiles.empty() ? 0 : 1) + (hp.tiles_global.empty() ? 0 : 1);
        // this device is through with the sample
        CUDA_OK(cudaEventRecord(cx.chunk_event(2 * S + 1 + smp), st));
      }
      if (M == 0) continue;
      // device 0, copy stream: once every device has flushed the sample, gather its coverage and send it home
      cx0.bind();
      for (uint32_t i = 0; i < n; ++i) CUDA_OK(cudaStreamWaitEvent(cs0, forests[i]->ctx->chunk_event(2 * S + 1 + smp), 0));
      CUDA_OK(pcs::launch_finalize(cs0, t_depth + smp * L, owner.d_row_locus.p, 1u, static_cast<uint32_t>(L),
                                   static_cast<uint32_t>(M), t_cov + smp * M));
      ++launches;
      // The tables go home in n_links row slices: device 0 forwards slice i to device i over NVLink (a few
      // microseconds) and device i's copy stream sends it to the host; slice 0 goes over device 0's own link.
      const char* dev_tbl[2] = {reinterpret_cast<const char*>(t_occ + smp * M), reinterpret_cast<const char*>(t_cov + smp * M)};
      char* host_tbl[2] = {reinterpret_cast<char*>(occ + smp * M), reinterpret_cast<char*>(cov + smp * M)};
      for (int t = 0; t < 2; ++t) {
        for (uint32_t i = 1; i < n_links; ++i) {  // the forwards first: they are short and the peers' links start early
          const size_t r_lo = slice_lo(i), nb = (slice_lo(i + 1) - r_lo) * sizeof(uint32_t);
          if (nb == 0) continue;
          uint32_t* scr = scratch[i] + (2 * smp + t) * slice_cap;
          CUDA_OK(cudaMemcpyPeerAsync(scr, forests[i]->ctx->device, dev_tbl[t] + r_lo * sizeof(uint32_t), cx0.device, nb, cs0));
          cudaEvent_t fw = cx0.chunk_event(3 * S + 2 + (2 * smp + t) * n_links + i);
          CUDA_OK(cudaEventRecord(fw, cs0));
          pcs_ctx& cx = *forests[i]->ctx;
          cx.bind();
          cudaStream_t cs = cx.copier();
          CUDA_OK(cudaStreamWaitEvent(cs, fw, 0));
          char* sp = stage + (2 * smp + t) * row_bytes + r_lo * sizeof(uint32_t);
          CUDA_OK(cudaMemcpyAsync(sp, scr, nb, cudaMemcpyDeviceToHost, cs));
          cudaEvent_t ev = cx.chunk_event(2 * smp + t);
          CUDA_OK(cudaEventRecord(ev, cs));
          chunks.push_back({host_tbl[t] + r_lo * sizeof(uint32_t), sp, nb});
          chunk_done.push_back(ev);
          cx0.bind();
        }
        const size_t nb0 = slice_lo(1) * sizeof(uint32_t);
        char* sp = stage + (2 * smp + t) * row_bytes;
        CUDA_OK(cudaMemcpyAsync(sp, dev_tbl[t], nb0, cudaMemcpyDeviceToHost, cs0));
        cudaEvent_t ev = cx0.chunk_event(2 * smp + t);
        CUDA_OK(cudaEventRecord(ev, cs0));
        chunks.push_back({host_tbl[t], sp, nb0});
        chunk_done.push_back(ev);
      }
    }
    lap("  plan + launch, all samples");
    cx0.bind();
    DevBuf<unsigned long long> d_sums;
    d_sums.alloc(4, cs0);
    CUDA_OK(cudaMemsetAsync(d_sums.p, 0, 4 * sizeof(unsigned long long), cs0));
    if (stats) {
      CUDA_OK(pcs::launch_sum_u32(cs0, t_depth, S * L, d_sums.p + 1));
      CUDA_OK(pcs::launch_sum_u32(cs0, t_occ, S * M, d_sums.p + 2));
      launches += (S * M != 0 ? 1 : 0) + (S * L != 0 ? 1 : 0);
    }
    unsigned long long* sums = cx0.counters_home();  // [0..2] checksums, [3] reads placed by this device
    CUDA_OK(cudaMemcpyAsync(sums, d_sums.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, cs0));
    std::vector<unsigned long long*> placed(n, nullptr);
    for (uint32_t i = 0; i < n; ++i) {  // every device's own read counter, into its context's pinned slot
      pcs_ctx& cx = *forests[i]->ctx;
      cx.bind();
      placed[i] = cx.counters_home() + 3;
      CUDA_OK(cudaMemcpyAsync(placed[i], d_counters[i].p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, cx.stream));
    }
    cx0.bind();
    drain_chunks_ev(cx0.device, chunks, chunk_done);
    for (uint32_t i = 0; i < n; ++i) {
      pcs_ctx& cx = *forests[i]->ctx;
      cx.bind();
      CUDA_OK(cudaStreamSynchronize(cx.stream));
      if (i && i < n_links) CUDA_OK(cudaStreamSynchronize(cx.copier()));
      d_counters[i].release();
      d_alias[i].release();
    }
    cx0.bind();
    CUDA_OK(cudaStreamSynchronize(cs0));
    d_sums.release();
    if (std::getenv("PCS_TIMING")) std::fprintf(stderr, "[pcs host]    %-28s %8.2f ms\n", "simulate_multi", now_ms() - t0);
    if (stats) {
      *stats = pcs_run_stats{};
      for (uint32_t i = 0; i < n; ++i) {
        forests[i]->ctx->bind();
        stats->kernel_ms = std::max(stats->kernel_ms, pipelined_kernel_ms(*forests[i]->ctx, S));
        stats->n_reads += *placed[i];
      }
      stats->kernel_launches = launches;
      stats->n_templates = templates;
      stats->sum_depth = sums[1];
      stats->sum_occurrences = sums[2];
      stats->h2d_bytes = h2d;
      stats->d2h_bytes = 2 * S * M * sizeof(uint32_t) + (4 + n) * sizeof(unsigned long long);
      stats->total_ms = now_ms() - t0;
    }
  });
}

int pcs_forest_set_reference(pcs_forest* fo, uint32_t chr, const char* bases, uint64_t len) {
  return guarded([&] {
    require(fo && bases, "bad arguments");
    const pcs::FlatForest& F = fo->host.flat;
    require(chr < F.n_chr, "chromosome out of range");
    require(len == F.chr_len[chr], "the reference sequence length differs from the chromosome length");
    fo->ref_chr.resize(F.n_chr);
    std::string& dst = fo->ref_chr[chr];
    dst.assign(bases, len);
    for (auto& ch : dst) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
    fo->seq_dirty = true;
  });
}

int pcs_forest_load_fasta(pcs_forest* fo, const char* path, const char* const* chr_names, uint32_t* n_loaded) {
  return guarded([&] {
    require(fo && path && chr_names, "bad arguments");
    const pcs::FlatForest& F = fo->host.flat;
    std::ifstream in(path);
    if (!in) throw std::runtime_error(std::string("The reference genome file \"") + path + "\" does not exists.");
    std::map<std::string, uint32_t> index;
    for (uint32_t c = 0; c < F.n_chr; ++c) index[chr_names[c]] = c;
    fo->ref_chr.resize(F.n_chr);
    std::string line, name, seq;
    uint32_t loaded = 0;
    auto flush = [&]() {
      if (name.empty()) return;
      std::string key = name;
      if (!index.count(key) && key.rfind("chr", 0) == 0) key = key.substr(3);
      auto it = index.find(key);
      if (it != index.end()) {
        require(seq.size() == F.chr_len[it->second], "a FASTA sequence length differs from the chromosome length");
        for (auto& ch : seq) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
        fo->ref_chr[it->second] = seq;
        ++loaded;
      }
    };
    while (std::getline(in, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (!line.empty() && line[0] == '>') {
        flush();
        size_t e = line.find_first_of(" \t", 1);
        name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
        seq.clear();
      } else {
        seq += line;
      }
    }
    flush();
    fo->seq_dirty = true;
    if (n_loaded) *n_loaded = loaded;
  });
}

int pcs_forest_set_alt(pcs_forest* fo, const uint32_t* alt_off, const char* alt_bytes) {
  return guarded([&] {
    require(fo && alt_off && alt_bytes, "bad arguments");
    const pcs::FlatForest& F = fo->host.flat;
    fo->alt_off.assign(alt_off, alt_off + F.n_mut + 1);
    require(fo->alt_off[0] == 0, "alt_off must start at 0");
    for (uint32_t m = 0; m < F.n_mut; ++m)
      require(fo->alt_off[m + 1] >= fo->alt_off[m] + 1, "every row needs a non-empty alt string");
    fo->alt_bytes.assign(alt_bytes, fo->alt_off[F.n_mut]);
    fo->seq_dirty = true;
  });
}

int pcs_plan_materialize(pcs_plan* pl, uint64_t cap, pcs_read_placement* placements, uint32_t* err_masks, uint8_t* seq,
                         uint8_t* qual, uint32_t* cigar, uint32_t* n_cigar, uint32_t* lengths, uint64_t* n_out) {
  return guarded([&] {
    require(pl && placements && seq && qual && cigar && n_cigar && lengths && n_out, "bad arguments");
    require(!err_masks || pl->host.info.read_size <= 32u * PCS_ERRMASK_WORDS,
            "error masks cover 256 read offsets: reads longer than that cannot be materialised with masks");
    std::vector<pcs::Tile> tiles = pl->host.tiles;
    tiles.insert(tiles.end(), pl->host.tiles_global.begin(), pl->host.tiles_global.end());
    Materialized m;
    materialize_tiles(*pl, tiles, cap, m);
    const uint32_t R = pl->host.info.read_size;
    *n_out = m.hdr.size();
    for (size_t i = 0; i < m.hdr.size(); ++i) {
      placement_of(pl->forest->host.flat, m.hdr[i], placements[i]);
      n_cigar[i] = m.hdr[i].n_cigar;
      lengths[i] = m.hdr[i].len;
      std::memcpy(cigar + i * pcs::kMaxCigar, m.hdr[i].cigar, sizeof(uint32_t) * pcs::kMaxCigar);
    }
    if (err_masks && !m.masks.empty()) std::memcpy(err_masks, m.masks.data(), m.masks.size() * sizeof(uint32_t));
    if (!m.seq.empty()) {
      std::memcpy(seq, m.seq.data(), m.hdr.size() * R);
      std::memcpy(qual, m.qual.data(), m.hdr.size() * R);
    }
  });
}

// One batch of materialised reads in flight: device records, their pinned landing area, the event that says they
// have landed.  The SAM writer keeps two: the GPU fills one while the host threads format the other.
struct SamBatch {
  DevBuf<pcs::Tile> d_tiles;
  DevBuf<pcs::SamHeader> d_hdr;
  DevBuf<uint8_t> d_seq, d_qual;
  DevBuf<unsigned long long> d_count;
  char* pinned = nullptr;  // [count (256 B) | hdr | seq | qual]
  size_t pinned_bytes = 0, cap = 0;
  cudaEvent_t landed = nullptr;
  uint32_t R = 0;
  ~SamBatch() {
    if (pinned) cudaFreeHost(pinned);
    if (landed) cudaEventDestroy(landed);
  }
  const unsigned long long* count() const { return reinterpret_cast<const unsigned long long*>(pinned); }
  const pcs::SamHeader* hdr() const { return reinterpret_cast<const pcs::SamHeader*>(pinned + 256); }
  const uint8_t* seq() const { return reinterpret_cast<const uint8_t*>(pinned + 256 + cap * sizeof(pcs::SamHeader)); }
  const uint8_t* qual() const { return seq() + cap * R; }
  // queue: tiles up, kernel, records + count back into pinned memory; nothing waits
  void launch(pcs_plan& pl, const pcs::SeqData& D, const std::vector<pcs::Tile>& tiles, uint64_t reads) {
    pcs_forest& fo = *pl.forest;
    cudaStream_t st = fo.ctx->stream;
    R = pl.host.info.read_size;
    if (!landed) CUDA_OK(cudaEventCreateWithFlags(&landed, cudaEventDisableTiming));
    cap = std::max<uint64_t>(reads, 1);
    const size_t need = 256 + cap * (sizeof(pcs::SamHeader) + 2 * static_cast<size_t>(R));
    if (need > pinned_bytes) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      pinned_bytes = 0;
      CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&pinned), need + need / 8, cudaHostAllocDefault));
      pinned_bytes = need + need / 8;
    }
    d_tiles.upload(tiles, st);
    d_hdr.alloc(cap, st);
    d_seq.alloc(cap * R, st);
    d_qual.alloc(cap * R, st);
    d_count.alloc(1, st);
    CUDA_OK(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), st));
    CUDA_OK(pcs::launch_materialize_tiles(st, d_tiles.p, static_cast<uint32_t>(tiles.size()), pl.d_entries.p, pl.d_entry_lo.p,
                                          fo.dev(), pl.host.model, D, d_hdr.p, nullptr, d_seq.p, d_qual.p, cap, d_count.p));
    CUDA_OK(cudaMemcpyAsync(pinned, d_count.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(pinned + 256, d_hdr.p, cap * sizeof(pcs::SamHeader), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(pinned + 256 + cap * sizeof(pcs::SamHeader), d_seq.p, cap * R, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(pinned + 256 + cap * (sizeof(pcs::SamHeader) + R), d_qual.p, cap * R, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaEventRecord(landed, st));
  }
};

int pcs_plan_write_sam(pcs_plan* pl, const pcs_sam_options* opt, uint64_t* n_written) {
  return guarded([&] {
    require(pl && opt && opt->output_dir && opt->chr_names && opt->sample_names, "bad arguments");
    namespace fs = std::filesystem;
    pcs_forest& fo = *pl->forest;
    const pcs::FlatForest& F = fo.host.flat;
    const fs::path dir(opt->output_dir);
    // ReadSimulator<>::Mode::CREATE refuses an existing directory, UPDATE adds files to it
    // (src/seq_simulation.cpp:545-549, vignettes/sequencing.Rmd:283-309)
    if (fs::exists(dir) && !opt->update)
      throw std::domain_error("The output directory \"" + dir.string() + "\" already exists: use update_SAM=TRUE to add files to it.");
    fs::create_directories(dir);
    const std::string fprefix = opt->filename_prefix ? opt->filename_prefix : "chr_";
    const std::string tprefix = opt->template_name_prefix ? opt->template_name_prefix : "r";
    const uint32_t R = pl->host.info.read_size, mates = pl->host.info.reads_per_template;
    std::vector<pcs::Tile> tiles = pl->host.tiles;
    tiles.insert(tiles.end(), pl->host.tiles_global.begin(), pl->host.tiles_global.end());
    for (const auto& t : tiles) require(fo.has_reference(t.chr), "the reference sequence of a sequenced chromosome is not loaded");
    fo.ctx->bind();
    const pcs::SeqData D = fo.seq_data();
    uint64_t written = 0, text_bytes = 0;
    double ms_wait = 0, ms_format = 0, ms_file = 0;  // PCS_TIMING: where the host's time of a SAM run goes
    // batches of ~256 k reads: two in flight (the GPU materialises batch k+1 while the host formats batch k)
    const uint64_t batch_cap = 1u << 18;
    const unsigned nt = std::max(1u, std::min(32u, host_threads()));
    SamBatch slots[2];
    for (uint32_t c = 0; c < F.n_chr; ++c) {
      std::vector<pcs::Tile> mine;
      for (const auto& t : tiles)
        if (t.chr == c) mine.push_back(t);
      if (mine.empty()) continue;
      std::stable_sort(mine.begin(), mine.end(), [](const pcs::Tile& a, const pcs::Tile& b) { return a.begin < b.begin; });
      fs::path file = dir / (fprefix + opt->chr_names[c] + ".sam");
      for (uint32_t k = 1; fs::exists(file); ++k) file = dir / (fprefix + opt->chr_names[c] + "_" + std::to_string(k) + ".sam");
      std::ofstream out(file, std::ios::binary);
      if (!out) throw std::runtime_error("cannot write \"" + file.string() + "\"");
      std::string head = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:" + std::string(opt->chr_names[c]) + "\tLN:" + std::to_string(F.chr_len[c]) + "\n";
      for (uint32_t s = 0; s < pl->host.info.n_out_samples; ++s)
        head += std::string("@RG\tID:") + opt->sample_names[s] + "\tSM:" + opt->sample_names[s] + "\tPL:ILLUMINA\n";
      head += "@PG\tID:pcs_seq\tPN:pcs_seq\n";
      out.write(head.data(), static_cast<std::streamsize>(head.size()));
      text_bytes += head.size();
      // the chromosome's batches
      std::vector<std::pair<size_t, size_t>> batches;  // [first tile, one past the last)
      std::vector<uint64_t> batch_reads;
      for (size_t i = 0; i < mine.size();) {
        const size_t first = i;
        uint64_t reads = 0;
        while (i < mine.size() && (i == first || reads + static_cast<uint64_t>(mine[i].n_templates) * mates <= batch_cap))
          reads += static_cast<uint64_t>(mine[i++].n_templates) * mates;
        batches.emplace_back(first, i);
        batch_reads.push_back(reads);
      }
      auto launch = [&](size_t k) {
        slots[k & 1].launch(*pl, D, std::vector<pcs::Tile>(mine.begin() + static_cast<std::ptrdiff_t>(batches[k].first),
                                                            mine.begin() + static_cast<std::ptrdiff_t>(batches[k].second)),
                            batch_reads[k]);
      };
      launch(0);
      // two sets of text buffers: a writer thread puts batch k into the file while the others format batch k+1
      std::vector<std::string> text_sets[2] = {std::vector<std::string>(nt), std::vector<std::string>(nt)};
      std::thread writer;
      struct JoinWriter {
        std::thread& t;
        ~JoinWriter() { if (t.joinable()) t.join(); }
      } join_writer{writer};
      for (size_t k = 0; k < batches.size(); ++k) {
        std::vector<std::string>& text = text_sets[k & 1];
        if (k + 1 < batches.size()) launch(k + 1);  // behind batch k on the stream: the GPU never waits for the host
        SamBatch& sb = slots[k & 1];
        double t_a = now_ms();
        CUDA_OK(cudaEventSynchronize(sb.landed));
        ms_wait += now_ms() - t_a;
        const uint64_t n = *sb.count();
        if (n > sb.cap) throw std::domain_error("internal: more reads materialised than planned");
        t_a = now_ms();
        std::vector<std::string> errors(nt);
        pcs::HostPool::get().run(nt, [&](size_t w) {
            std::string& s = text[w];
            s.clear();
            const uint64_t lo = n * w / nt, hi = n * (w + 1) / nt;
            s.reserve((hi - lo) * (2 * static_cast<size_t>(R) + 96));
            for (uint64_t r = lo; r < hi; ++r) {
              const pcs::SamHeader& h = sb.hdr()[r];
              if (h.flags & 4u) {  // a CIGAR the record cannot hold would no longer describe SEQ: refuse, never truncate
                errors[w] = "a read carries more indels than a CIGAR of " + std::to_string(pcs::kMaxCigar) + " operations can describe";
                return;
              }
              append_sam_line(s, h, sb.seq() + r * R, sb.qual() + r * R, tprefix, opt->chr_names[c], opt->sample_names[h.chr_sample >> 16]);
            }
          }, nt);
        for (const auto& e : errors)
          if (!e.empty()) throw std::domain_error(e);
        ms_format += now_ms() - t_a;
        t_a = now_ms();
        if (writer.joinable()) writer.join();  // batch k-1 is in the file (and its buffers are free for batch k+1)
        ms_file += now_ms() - t_a;
        for (const auto& s : text) text_bytes += s.size();
        writer = std::thread([&out, &text] {
          for (const auto& s : text) out.write(s.data(), static_cast<std::streamsize>(s.size()));
        });
        written += n;
      }
      {
        const double t_a = now_ms();
        if (writer.joinable()) writer.join();
        out.flush();
        ms_file += now_ms() - t_a;
      }
      if (!out) throw std::runtime_error("writing \"" + file.string() + "\" failed");
    }
    CUDA_OK(cudaStreamSynchronize(fo.ctx->stream));
    if (std::getenv("PCS_TIMING"))
      std::fprintf(stderr, "[pcs sam] reads %llu text_bytes %llu gpu_materialise_ms %.1f host_format_ms %.1f file_write_ms %.1f\n",
                   static_cast<unsigned long long>(written), static_cast<unsigned long long>(text_bytes), ms_wait, ms_format, ms_file);
    if (n_written) *n_written = written;
  });
}

int pcs_plan_coverage_track(pcs_plan* pl, uint32_t bin_bp, uint64_t* chr_bin_off, uint32_t* track, uint64_t cap_bins,
                            uint64_t* n_bins_out) {
  return guarded([&] {
    require(pl && chr_bin_off && n_bins_out, "bad arguments");
    require(bin_bp >= 64 && (bin_bp & (bin_bp - 1)) == 0, "the bin size must be a power of two, at least 64");
    uint32_t shift = 0;
    while ((1u << shift) < bin_bp) ++shift;
    pcs_forest& fo = *pl->forest;
    pcs_ctx& cx = *fo.ctx;
    const pcs::FlatForest& F = fo.host.flat;
    std::vector<uint64_t> off(F.n_chr + 1, 0);
    for (uint32_t c = 0; c < F.n_chr; ++c) off[c + 1] = off[c] + (static_cast<uint64_t>(F.chr_len[c]) >> shift) + 1;
    const uint64_t n_bins = off[F.n_chr];
    std::copy(off.begin(), off.end(), chr_bin_off);
    *n_bins_out = n_bins;
    if (!track) return;  // sizes only
    const size_t S = pl->host.info.n_out_samples;
    require(cap_bins >= n_bins, "track capacity too small");
    cx.bind();
    cudaStream_t st = cx.stream;
    DevBuf<uint64_t> d_off;
    DevBuf<uint32_t> d_track;
    d_off.upload(off, st);
    d_track.alloc(S * n_bins, st);
    CUDA_OK(cudaMemsetAsync(d_track.p, 0, d_track.bytes(), st));
    uint32_t max_len = 0;
    for (const auto* v : {&pl->host.tiles, &pl->host.tiles_global})
      for (const auto& t : *v) max_len = std::max(max_len, t.len);
    CUDA_OK(pcs::launch_coverage_track(st, pl->d_tiles.p, static_cast<uint32_t>(pl->host.tiles.size()), pl->d_entries.p,
                                       fo.dev(), pl->host.model, shift, max_len, d_off.p, n_bins, d_track.p));
    CUDA_OK(pcs::launch_coverage_track(st, pl->d_tiles_global.p, static_cast<uint32_t>(pl->host.tiles_global.size()),
                                       pl->d_entries.p, fo.dev(), pl->host.model, shift, max_len, d_off.p, n_bins, d_track.p));
    copy_out(cx, st, {{track, d_track.p, d_track.bytes()}});
  });
}

int pcs_memcpy_d2h(pcs_ctx* cx, void* host_dst, const void* dev_src, size_t bytes) {
  return guarded([&] {
    require(cx && (bytes == 0 || (host_dst && dev_src)), "bad arguments");
    cx->bind();
    if (bytes) {
      CUDA_OK(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, cx->stream));
      CUDA_OK(cudaStreamSynchronize(cx->stream));
    }
  });
}

int pcs_plan_trace(pcs_plan* pl, pcs_read_placement* rec, uint32_t* masks, uint64_t cap, uint64_t* n_out) {
  return guarded([&] {
    require(pl && rec && n_out, "bad arguments");
    require(!masks || pl->host.model.sequencer == PCS_SEQ_ERRORLESS || pl->host.info.read_size <= 32u * PCS_ERRMASK_WORDS,
            "error masks cover 256 read offsets: reads longer than that cannot be traced with masks");
    pcs_forest& fo = *pl->forest;
    pcs_ctx& cx = *fo.ctx;
    cx.bind();
    cudaStream_t st = cx.stream;
    DevBuf<pcs::DevPlacement> d_rec;
    DevBuf<uint32_t> d_masks;
    d_rec.alloc(cap, st);
    if (masks) d_masks.alloc(cap * PCS_ERRMASK_WORDS, st);
    CUDA_OK(cudaMemsetAsync(pl->d_counters.p, 0, 4 * sizeof(unsigned long long), st));
    CUDA_OK(pcs::launch_trace_tiles(st, pl->d_tiles.p, static_cast<uint32_t>(pl->host.tiles.size()), pl->d_entries.p,
                                    pl->d_entry_lo.p, fo.dev(), pl->host.model, pl->d_counters.p, d_rec.p, d_masks.p, cap,
                                    pl->d_counters.p + 3));
    CUDA_OK(pcs::launch_trace_tiles(st, pl->d_tiles_global.p, static_cast<uint32_t>(pl->host.tiles_global.size()),
                                    pl->d_entries.p, pl->d_entry_lo.p, fo.dev(), pl->host.model, pl->d_counters.p, d_rec.p, d_masks.p,
                                    cap, pl->d_counters.p + 3));
    unsigned long long counters[4];
    CUDA_OK(cudaMemcpyAsync(counters, pl->d_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    *n_out = counters[3];
    if (counters[3] > cap) throw std::domain_error("trace capacity too small");
    std::vector<pcs::DevPlacement> h(counters[3]);
    if (!h.empty()) CUDA_OK(cudaMemcpy(h.data(), d_rec.p, h.size() * sizeof(pcs::DevPlacement), cudaMemcpyDeviceToHost));
    if (masks && !h.empty())
      CUDA_OK(cudaMemcpy(masks, d_masks.p, h.size() * PCS_ERRMASK_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); ++i) {
      uint32_t chr = h[i].chr_sample & 0xffffu;
      const pcs::HapRec& hr = fo.host.flat.chr_haps[chr][h[i].hap];
      rec[i].cell = hr.cell;
      rec[i].start = h[i].start;
      rec[i].chr = static_cast<uint16_t>(chr);
      rec[i].allele = hr.allele;
      rec[i].sample = static_cast<uint16_t>(h[i].chr_sample >> 16);
      rec[i].flags = hr.kind == pcs::HAP_TUMOUR ? PCS_PLACE_TUMOUR
                     : hr.kind == pcs::HAP_NORMAL_PLAIN ? PCS_PLACE_NORMAL_PLAIN : PCS_PLACE_NORMAL_PRENEO;
    }
  });
}

int pcs_simulate(pcs_forest* fo, const pcs_seq_params* params, uint32_t* occ, uint32_t* cov, pcs_run_stats* stats) {
  if (std::getenv("PCS_NO_PIPELINE") == nullptr)
    return guarded([&] {
      require(fo && params, "bad arguments");
      validate(*params);
      simulate_tables(*fo, *params, occ, cov, stats);
    });
  pcs_plan* pl = nullptr;
  int rc = pcs_plan_create(fo, params, &pl);
  if (rc != PCS_OK) return rc;
  rc = pcs_plan_run(pl, PCS_RUN_HOST_OUTPUT, occ, cov, stats);
  if (rc == PCS_OK && stats) stats->h2d_bytes = pl->h2d_bytes;
  std::string keep = g_err;
  pcs_plan_free(pl);
  g_err = keep;
  return rc;
}

int pcs_count_injected(pcs_forest* fo, uint32_t n_out_samples, uint32_t read_size, const pcs_read_placement* rec,
                       const uint32_t* masks, uint64_t n, uint32_t* occ, uint32_t* cov, pcs_run_stats* stats) {
  return guarded([&] {
    require(fo && occ && cov && (rec || n == 0), "bad arguments");
    require(read_size >= 1 && read_size <= 65535, "read_size must be in [1, 65535]");
    require(n_out_samples >= 1 && n_out_samples <= 65535, "n_out_samples out of range");
    require(!masks || read_size <= 32u * PCS_ERRMASK_WORDS,
            "error masks cover 256 read offsets: reads longer than that cannot be injected with masks");
    const pcs::FlatForest& F = fo->host.flat;
    fo->host.build_lookup();
    std::vector<pcs::DevPlacement> h(n);
    for (uint64_t i = 0; i < n; ++i) {
      const pcs_read_placement& r = rec[i];
      require(r.chr < F.n_chr, "placement chromosome out of range");
      require(r.sample < n_out_samples, "placement sample out of range");
      require(r.flags <= PCS_PLACE_NORMAL_PRENEO, "unknown placement flags");
      uint8_t kind = r.flags == PCS_PLACE_TUMOUR ? pcs::HAP_TUMOUR
                     : r.flags == PCS_PLACE_NORMAL_PLAIN ? pcs::HAP_NORMAL_PLAIN : pcs::HAP_NORMAL_PRENEO;
      uint64_t key = HostForest::hap_key(kind, r.cell, r.allele);
      const auto& lk = fo->host.lookup[r.chr];
      auto it = std::lower_bound(lk.begin(), lk.end(), std::make_pair(key, 0u));
      require(it != lk.end() && it->first == key, "placement names a missing allele");
      const pcs::HapRec& hr = F.chr_haps[r.chr][it->second];
      uint32_t frag_end = 0;
      for (const auto& fr : F.fragsets[hr.fragset])
        if (r.start >= fr.b && r.start <= fr.e) frag_end = fr.e;
      require(frag_end != 0, "placement starts outside every fragment of the allele");
      h[i] = pcs::DevPlacement{it->second, r.start, frag_end, static_cast<uint32_t>(r.chr) | (static_cast<uint32_t>(r.sample) << 16)};
    }
    pcs_ctx& cx = *fo->ctx;
    cx.bind();
    cudaStream_t st = cx.stream;
    const double t0 = now_ms();
    const size_t S = n_out_samples, M = F.n_mut, L = F.locus_pos.size();
    DevBuf<pcs::DevPlacement> d_rec;
    DevBuf<uint32_t> d_masks, d_depth, d_occ, d_cov;
    DevBuf<unsigned long long> d_cnt;
    uint64_t h2d = d_rec.upload(h, st);
    if (masks) {
      d_masks.alloc(n * PCS_ERRMASK_WORDS, st);
      if (n) CUDA_OK(cudaMemcpyAsync(d_masks.p, masks, d_masks.bytes(), cudaMemcpyHostToDevice, st));
      h2d += d_masks.bytes();
    }
    d_depth.alloc(S * L, st);
    d_occ.alloc(S * M, st);
    d_cov.alloc(S * M, st);
    d_cnt.alloc(4, st);
    if (S * L != 0) CUDA_OK(cudaMemsetAsync(d_depth.p, 0, d_depth.bytes(), st));
    if (S * M != 0) CUDA_OK(cudaMemsetAsync(d_occ.p, 0, d_occ.bytes(), st));
    CUDA_OK(cudaMemsetAsync(d_cnt.p, 0, d_cnt.bytes(), st));
    CUDA_OK(cudaEventRecord(cx.ev[1], st));
    CUDA_OK(pcs::launch_count_injected(st, d_rec.p, masks ? d_masks.p : nullptr, n, fo->dev(), read_size, d_depth.p, d_occ.p));
    CUDA_OK(cudaEventRecord(cx.ev[2], st));
    CUDA_OK(pcs::launch_finalize(st, d_depth.p, fo->d_row_locus.p, static_cast<uint32_t>(S), static_cast<uint32_t>(L),
                                 static_cast<uint32_t>(M), d_cov.p));
    CUDA_OK(pcs::launch_sum_u32(st, d_depth.p, S * L, d_cnt.p + 1));
    CUDA_OK(pcs::launch_sum_u32(st, d_occ.p, S * M, d_cnt.p + 2));
    if (S * M != 0) {
      CUDA_OK(cudaMemcpyAsync(occ, d_occ.p, d_occ.bytes(), cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaMemcpyAsync(cov, d_cov.p, d_cov.bytes(), cudaMemcpyDeviceToHost, st));
    }
    unsigned long long counters[4] = {0, 0, 0, 0};
    CUDA_OK(cudaMemcpyAsync(counters, d_cnt.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (stats) {
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, cx.ev[1], cx.ev[2]));
      stats->kernel_ms = ms;
      stats->total_ms = now_ms() - t0;
      stats->kernel_launches = (n ? 1 : 0) + (S * M != 0 ? 2 : 0) + (S * L != 0 ? 1 : 0);
      stats->n_templates = n;
      stats->n_reads = n;
      stats->sum_depth = counters[1];
      stats->sum_occurrences = counters[2];
      stats->h2d_bytes = h2d;
      stats->d2h_bytes = 2 * S * M * sizeof(uint32_t) + sizeof(counters);
    }
  });
}

int pcs_active_rows(pcs_forest* fo, const uint32_t* occ, uint32_t n_out_samples, int include_non_sequenced,
                    const pcs_seq_params* params, uint32_t* rows_out, uint32_t* n_rows) {
  return guarded([&] {
    require(fo && occ && rows_out && n_rows, "bad arguments");
    const pcs::FlatForest& F = fo->host.flat;
    std::vector<uint8_t> carried;
    if (include_non_sequenced) {
      fo->ensure_host_instances();
      carried = carried_rows(fo->host, params);
    }
    uint32_t k = 0;
    for (uint32_t m = 0; m < F.n_mut; ++m) {
      bool on = include_non_sequenced && carried[m];
      for (uint32_t s = 0; s < n_out_samples && !on; ++s) on = occ[static_cast<size_t>(s) * F.n_mut + m] > 0;
      if (on) rows_out[k++] = m;
    }
    *n_rows = k;
  });
}

int pcs_simulate_result(pcs_forest* fo, const pcs_seq_params* params, int include_non_sequenced, int with_vaf,
                        pcs_result** out, pcs_run_stats* stats) {
  return guarded([&] {
    require(fo && params && out, "bad arguments");
    validate(*params);
    pcs_ctx& cx = *fo->ctx;
    cx.bind();
    cudaStream_t st = cx.stream;
    const double t0 = now_ms();
    CallTables T;
    pcs_run_stats rs{};
    run_pipelined(*fo, *params, false, false, T, rs, [](uint32_t, uint32_t, uint32_t, uint64_t) {});
    uint64_t extra = 0;
    if (stats) {
      CUDA_OK(pcs::launch_sum_u32(st, T.depth.p, T.S * T.L, T.counters.p + 1));
      CUDA_OK(pcs::launch_sum_u32(st, T.occ.p, T.S * T.M, T.counters.p + 2));
      extra += (T.S * T.M != 0 ? 1 : 0) + (T.S * T.L != 0 ? 1 : 0);
    }
    unsigned long long* counters = cx.counters_home();
    CUDA_OK(cudaMemcpyAsync(counters, T.counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    std::unique_ptr<pcs_result> res = assemble_result(*fo, T.occ.p, T.depth.p, T.S, include_non_sequenced, params, with_vaf != 0);
    extra += T.S * T.M != 0 ? 3 : 0;
    CUDA_OK(cudaStreamSynchronize(st));
    rs.kernel_ms = pipelined_kernel_ms(cx, T.S);
    rs.total_ms = now_ms() - t0;
    rs.kernel_launches += extra;
    rs.n_reads = counters[0];
    rs.sum_depth = counters[1];
    rs.sum_occurrences = counters[2];
    rs.d2h_bytes = 4 * sizeof(unsigned long long) + sizeof(uint32_t);
    res->stats = rs;
    if (stats) *stats = rs;
    if (std::getenv("PCS_TIMING")) std::fprintf(stderr, "[pcs host]    %-28s %8.2f ms\n", "simulate_result (to compact)", rs.total_ms);
    *out = res.release();
  });
}

int pcs_plan_result(pcs_plan* pl, int include_non_sequenced, const pcs_seq_params* params, int with_vaf, pcs_result** out) {
  return guarded([&] {
    require(pl && out, "bad arguments");
    require(pl->last_occ != nullptr || static_cast<size_t>(pl->host.info.n_out_samples) * pl->host.info.n_mut == 0,
            "the plan has not been run (pcs_plan_run)");
    std::unique_ptr<pcs_result> res = assemble_result(*pl->forest, pl->last_occ, pl->d_depth.p, pl->host.info.n_out_samples,
                                                      include_non_sequenced, params, with_vaf != 0);
    CUDA_OK(cudaStreamSynchronize(pl->forest->ctx->stream));
    *out = res.release();
  });
}

int pcs_result_info(const pcs_result* res, uint32_t* n_rows, uint32_t* n_samples, int* has_vaf) {
  return guarded([&] {
    require(res != nullptr, "result is NULL");
    if (n_rows) *n_rows = res->n_rows;
    if (n_samples) *n_samples = res->n_samples;
    if (has_vaf) *has_vaf = res->d_vaf.p != nullptr || res->n_rows == 0;
  });
}

int pcs_result_fetch(pcs_result* res, uint32_t* rows, int32_t* const* occ_cols, int32_t* const* cov_cols,
                     double* const* vaf_cols, uint64_t* d2h_bytes) {
  return guarded([&] {
    require(res != nullptr, "result is NULL");
    const size_t n = res->n_rows, S = res->n_samples;
    std::vector<OutCopy> items;
    if (n) {
      if (rows) items.push_back({rows, res->d_rows.p, n * sizeof(uint32_t)});
      for (size_t sm = 0; sm < S; ++sm) {
        if (occ_cols && occ_cols[sm]) items.push_back({occ_cols[sm], res->d_occ.p + sm * n, n * sizeof(uint32_t)});
        if (cov_cols && cov_cols[sm]) items.push_back({cov_cols[sm], res->d_cov.p + sm * n, n * sizeof(uint32_t)});
        if (vaf_cols && vaf_cols[sm]) {
          require(res->d_vaf.p != nullptr, "the result was assembled without VAF columns");
          items.push_back({vaf_cols[sm], res->d_vaf.p + sm * n, n * sizeof(double)});
        }
      }
    }
    pcs_ctx& cx = *res->forest->ctx;
    cx.bind();
    const double t0 = now_ms();
    const uint64_t b = copy_out(cx, cx.stream, items);
    if (std::getenv("PCS_TIMING"))
      std::fprintf(stderr, "[pcs host]    %-28s %8.2f ms (%.1f MB)\n", "result fetch", now_ms() - t0, b / 1e6);
    if (d2h_bytes) *d2h_bytes = b;
  });
}

int pcs_result_free(pcs_result* res) {
  return guarded([&] {
    if (!res) return;
    cudaSetDevice(res->forest->ctx->device);
    delete res;
  });
}

// ------------------------------------------------ host-side column builders
// The annotation columns of the data frame (chr, ref, alt, causes, classes: src/seq_simulation.cpp:52-90) draw
// millions of rows from a handful of distinct strings.  The caller keeps one small code per ROW OF THE FOREST and
// the table of distinct strings; these build the column of the ACTIVE rows in one parallel pass, straight into
// the layout an Arrow large_string array (pandas) wraps without a copy.  (An R shim does the same with one
// mkChar per distinct string and one SET_STRING_ELT per row.)
int pcs_host_gather(const uint32_t* rows, uint64_t n, const void* src, uint32_t elem_bytes, void* dst) {
  return guarded([&] {
    require((rows && src && dst) || n == 0, "bad arguments");
    require(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, "element size must be 1, 2, 4 or 8");
    const size_t n_chunks = std::max<size_t>(1, std::min<size_t>(64, n >> 16));
    host_tasks(n_chunks, [&](size_t k) {
      const uint64_t lo = n * k / n_chunks, hi = n * (k + 1) / n_chunks;
      switch (elem_bytes) {
        case 1: for (uint64_t i = lo; i < hi; ++i) static_cast<uint8_t*>(dst)[i] = static_cast<const uint8_t*>(src)[rows[i]]; break;
        case 2: for (uint64_t i = lo; i < hi; ++i) static_cast<uint16_t*>(dst)[i] = static_cast<const uint16_t*>(src)[rows[i]]; break;
        case 4: for (uint64_t i = lo; i < hi; ++i) static_cast<uint32_t*>(dst)[i] = static_cast<const uint32_t*>(src)[rows[i]]; break;
        default: for (uint64_t i = lo; i < hi; ++i) static_cast<uint64_t*>(dst)[i] = static_cast<const uint64_t*>(src)[rows[i]]; break;
      }
    });
  });
}

int pcs_host_string_column(const uint32_t* rows, uint64_t n, const uint16_t* codes, const char* const* table,
                           uint32_t n_table, int64_t* offsets, char* data, uint64_t data_cap, uint8_t* validity,
                           uint64_t* data_len, uint64_t* null_count) {
  return guarded([&] {
    require(((rows && codes) || n == 0) && table && offsets && data_len, "bad arguments");
    std::vector<uint32_t> len(n_table, 0);
    for (uint32_t t = 0; t < n_table; ++t) len[t] = table[t] ? static_cast<uint32_t>(std::strlen(table[t])) : 0u;
    // chunks of a multiple of 8 rows: no two threads share a validity byte
    const size_t n_chunks = std::max<size_t>(1, std::min<size_t>(64, n >> 16));
    auto chunk_lo = [&](size_t k) { return k == n_chunks ? n : (n * k / n_chunks) & ~static_cast<uint64_t>(7); };
    std::vector<uint64_t> bytes(n_chunks + 1, 0), nulls(n_chunks, 0);
    host_tasks(n_chunks, [&](size_t k) {
      uint64_t b = 0;
      for (uint64_t i = chunk_lo(k); i < chunk_lo(k + 1); ++i) {
        const uint32_t c = codes[rows[i]];
        if (c >= n_table) throw std::domain_error("string code outside the table");
        b += len[c];
      }
      bytes[k + 1] = b;
    });
    for (size_t k = 0; k < n_chunks; ++k) bytes[k + 1] += bytes[k];
    *data_len = bytes[n_chunks];
    offsets[n] = static_cast<int64_t>(bytes[n_chunks]);
    const bool fill = data != nullptr;
    require(!fill || data_cap >= bytes[n_chunks], "data buffer too small");
    host_tasks(n_chunks, [&](size_t k) {
      uint64_t at = bytes[k], nn = 0;
      const uint64_t lo = chunk_lo(k), hi = chunk_lo(k + 1);
      if (validity) std::memset(validity + lo / 8, 0, (hi - lo + 7) / 8);
      for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t c = codes[rows[i]];
        offsets[i] = static_cast<int64_t>(at);
        if (table[c]) {
          if (fill) std::memcpy(data + at, table[c], len[c]);
          if (validity) validity[i >> 3] |= static_cast<uint8_t>(1u << (i & 7));
          at += len[c];
        } else {
          ++nn;
        }
      }
      nulls[k] = nn;
    });
    if (null_count) {
      *null_count = 0;
      for (uint64_t x : nulls) *null_count += x;
    }
  });
}

// ------------------------------------------------ host-only introspection
// No GPU needed: tests use these to check the flattened view against explicit
// per-cell genomes, and the planner's shard partition at world_size > 1.

int pcs_flat_create(const pcs_forest_desc* desc, pcs_flat** out) {
  return guarded([&] {
    require(desc && out, "bad arguments");
    auto fl = std::make_unique<pcs_flat>();
    // same code path as pcs_forest_upload: the tables are written into a lent block (plain memory here)
    const size_t bytes = pcs::flat_store_bytes(*desc);
    fl->block.reset(new char[bytes]);
    fl->host.flat.store.base = fl->block.get();
    fl->host.flat.store.capacity = bytes;
    Lap lap;
    pcs::flatten_forest(*desc, fl->host.flat, host_threads());
    lap("flatten_forest");
    require(fl->host.flat.store.heap.empty(), "internal: the lent block was too small for the flat tables");
    fl->host.build_groups(fl->host.flat.leaf_sample.data(), fl->host.flat.n_samples);
    lap("sample groups");
    *out = fl.release();
  });
}

int pcs_flat_create_genomes(const pcs_cell_genomes_desc* desc, pcs_flat** out) {
  return guarded([&] {
    require(desc && out, "bad arguments");
    auto fl = std::make_unique<pcs_flat>();
    const size_t bytes = pcs::flat_store_bytes(*desc);
    fl->block.reset(new char[bytes]);
    fl->host.flat.store.base = fl->block.get();
    fl->host.flat.store.capacity = bytes;
    pcs::flatten_cell_genomes(*desc, fl->host.flat, host_threads());
    require(fl->host.flat.store.heap.empty(), "internal: the lent block was too small for the flat tables");
    fl->host.build_groups(fl->host.flat.leaf_sample.data(), fl->host.flat.n_samples);
    *out = fl.release();
  });
}

int pcs_flat_free(pcs_flat* fl) {
  delete fl;
  return PCS_OK;
}

int pcs_flat_set_groups(pcs_flat* fl, const uint32_t* leaf_group, uint32_t n_groups) {
  return guarded([&] {
    require(fl != nullptr, "flat is NULL");
    if (leaf_group)
      fl->host.build_groups(leaf_group, n_groups);
    else
      fl->host.build_groups(fl->host.flat.leaf_sample.data(), fl->host.flat.n_samples);
  });
}

int pcs_flat_info(const pcs_flat* fl, uint64_t out[6]) {
  return guarded([&] {
    require(fl && out, "bad arguments");
    const pcs::FlatForest& F = fl->host.flat;
    uint64_t haps = 0;
    for (const auto& v : F.chr_haps) haps += v.size();
    out[0] = F.locus_pos.size();
    out[1] = F.n_inst;
    out[2] = haps;
    out[3] = F.fragsets.size();
    out[4] = F.pieces.size();
    out[5] = 0;
  });
}

int pcs_flat_cell_haps(pcs_flat* fl, uint32_t kind, uint32_t cell, uint32_t chr, uint32_t cap, uint16_t* allele,
                       uint32_t* hap, uint32_t* fragset, uint32_t* n) {
  return guarded([&] {
    require(fl && n, "bad arguments");
    const pcs::FlatForest& F = fl->host.flat;
    require(chr < F.n_chr, "chromosome out of range");
    fl->host.build_lookup();
    const auto& lk = fl->host.lookup[chr];
    auto it = std::lower_bound(lk.begin(), lk.end(), std::make_pair(HostForest::hap_key(static_cast<uint8_t>(kind), cell, 0), 0u));
    uint32_t k = 0;
    for (; it != lk.end() && (it->first >> 16) == (HostForest::hap_key(static_cast<uint8_t>(kind), cell, 0) >> 16); ++it) {
      if (k < cap) {
        allele[k] = static_cast<uint16_t>(it->first & 0xffffu);
        hap[k] = it->second;
        fragset[k] = F.chr_haps[chr][it->second].fragset;
      }
      ++k;
    }
    *n = k;
  });
}

int pcs_flat_fragset(const pcs_flat* fl, uint32_t fragset, uint32_t cap, uint32_t* begin, uint32_t* end, uint32_t* n) {
  return guarded([&] {
    require(fl && n, "bad arguments");
    const pcs::FlatForest& F = fl->host.flat;
    require(fragset < F.fragsets.size(), "fragment set out of range");
    uint32_t k = 0;
    for (const auto& fr : F.fragsets[fragset]) {
      if (k < cap) {
        begin[k] = fr.b;
        end[k] = fr.e;
      }
      ++k;
    }
    *n = k;
  });
}

int pcs_flat_hap_rows(const pcs_flat* fl, uint32_t chr, uint32_t hap, uint32_t cap, uint32_t* rows, uint32_t* n) {
  return guarded([&] {
    require(fl && n, "bad arguments");
    const pcs::FlatForest& F = fl->host.flat;
    require(chr < F.n_chr, "chromosome out of range");
    uint32_t k = 0;
    for (uint32_t l = F.chr_locus_off[chr]; l < F.chr_locus_off[chr + 1]; ++l)
      for (uint32_t i = F.locus_inst_off[l]; i < F.locus_inst_off[l + 1]; ++i)
        if (hap - F.inst[i].lo < F.inst[i].span) {
          if (k < cap) rows[k] = F.inst[i].row;
          ++k;
        }
    *n = k;
  });
}

int pcs_flat_group_list(const pcs_flat* fl, uint32_t group, uint32_t fragset, uint32_t cap, uint32_t* haps,
                        uint32_t* offset, uint32_t* n) {
  return guarded([&] {
    require(fl && n, "bad arguments");
    const HostForest& H = fl->host;
    require(group < H.n_groups + 2 && fragset < H.flat.fragsets.size(), "group or fragment set out of range");
    auto it = H.list_index.find(HostForest::list_key(group, fragset));
    *n = 0;
    if (offset) *offset = 0;
    if (it == H.list_index.end()) return;
    *n = it->second.second;
    if (offset) *offset = it->second.first;
    require(static_cast<size_t>(it->second.first) + it->second.second <= H.hap_list.size(), "internal: list outside hap_list");
    for (uint32_t i = 0; i < it->second.second && i < cap; ++i) haps[i] = H.hap_list[it->second.first + i];
  });
}

int pcs_flat_plan(const pcs_flat* fl, const pcs_seq_params* params, pcs_plan_info* info, uint64_t cap,
                  uint32_t* tile_id, uint32_t* tile_templates, uint32_t* tile_sample, uint32_t* tile_chr,
                  uint32_t* tile_begin, uint32_t* tile_len) {
  return guarded([&] {
    require(fl && params && info, "bad arguments");
    validate(*params);
    HostPlan pl = make_host_plan(fl->host, *params);
    *info = pl.info;
    std::vector<pcs::Tile> both = pl.tiles;
    both.insert(both.end(), pl.tiles_global.begin(), pl.tiles_global.end());
    std::stable_sort(both.begin(), both.end(), [](const pcs::Tile& a, const pcs::Tile& b) { return a.n_templates > b.n_templates; });
    for (size_t i = 0; i < both.size() && i < cap; ++i) {
      if (tile_id) tile_id[i] = both[i].id;
      if (tile_templates) tile_templates[i] = both[i].n_templates;
      if (tile_sample) tile_sample[i] = both[i].sample;
      if (tile_chr) tile_chr[i] = both[i].chr;
      if (tile_begin) tile_begin[i] = both[i].begin;
      if (tile_len) tile_len[i] = both[i].len;
    }
  });
}


int pcs_flat_plan_thinning(const pcs_flat* fl, const pcs_seq_params* params, uint64_t cap, uint32_t* tile_id,
                           uint32_t* thin, uint32_t* u_len, uint32_t* tail_off, uint32_t* n_useful) {
  return guarded([&] {
    require(fl && params, "bad arguments");
    validate(*params);
    const HostPlan pl = make_host_plan(fl->host, *params);
    size_t i = 0;
    for (const auto* v : {&pl.tiles, &pl.tiles_global})
      for (const pcs::Tile& t : *v) {
        if (i >= cap) return;
        if (tile_id) tile_id[i] = t.id;
        if (thin) thin[i] = t.thin;
        if (u_len) u_len[i] = t.u_len;
        if (tail_off) tail_off[i] = t.tail_off;
        if (n_useful) n_useful[i] = t.n_useful;
        ++i;
      }
  });
}

int pcs_host_binomial(uint32_t seed, uint64_t n, double p, uint64_t count, uint64_t* out) {
  return guarded([&] {
    require(out || count == 0, "bad arguments");
    require(n < (1ull << 53) && p >= 0.0 && p <= 1.0, "Binomial(n, p): n < 2^53 and 0 <= p <= 1");
    pcs::PlanRng rng(seed, 0x7e57u, 0u, 0u, 0u);
    for (uint64_t i = 0; i < count; ++i) out[i] = pcs::binomial(rng, n, p);
  });
}

int pcs_flat_instances(const pcs_flat* fl, uint32_t* inst, uint32_t* locus_inst_off) {
  return guarded([&] {
    require(fl && inst && locus_inst_off, "bad arguments");
    const pcs::FlatForest& F = fl->host.flat;
    std::memcpy(inst, F.inst.data(), F.inst.size() * sizeof(pcs::Inst));
    std::memcpy(locus_inst_off, F.locus_inst_off.data(), F.locus_inst_off.size() * sizeof(uint32_t));
  });
}

int pcs_flat_hap_list(const pcs_flat* fl, uint32_t offset, uint32_t n, uint32_t* haps) {
  return guarded([&] {
    require(fl && (n == 0 || haps), "bad arguments");
    require(static_cast<size_t>(offset) + n <= fl->host.hap_list.size(), "slice outside hap_list");
    std::copy_n(fl->host.hap_list.begin() + offset, n, haps);
  });
}

int pcs_flat_tile_entries(const pcs_flat* fl, const pcs_seq_params* params, uint32_t tile_id, uint32_t cap,
                          uint32_t* thr, uint32_t* list_off, uint32_t* list_n, uint32_t* frag_end, uint32_t* n) {
  return guarded([&] {
    require(fl && params && n, "bad arguments");
    validate(*params);
    pcs_seq_params P = *params;
    P.shard_rank = 0;
    P.shard_count = 1;
    const HostPlan pl = make_host_plan(fl->host, P);
    std::unordered_map<uint32_t, uint32_t> n_of;  // list offset -> haplotypes in the list
    for (const auto& kv : fl->host.list_index) n_of[kv.second.first] = kv.second.second;
    for (const auto* v : {&pl.tiles, &pl.tiles_global})
      for (const pcs::Tile& t : *v) {
        if (t.id != tile_id) continue;
        *n = t.n_entries;
        for (uint32_t e = 0; e < t.n_entries && e < cap; ++e) {
          const pcs::Entry& en = pl.entries[t.entry_off + e];
          if (thr) thr[e] = en.thr;
          if (list_off) list_off[e] = en.list_off;
          if (list_n) list_n[e] = n_of.at(en.list_off);
          if (frag_end) frag_end[e] = en.frag_end;
        }
        return;
      }
    throw std::domain_error("no tile with that id holds templates");
  });
}

int pcs_flat_draw(const pcs_flat* fl, const pcs_seq_params* params, uint32_t tile_id, uint64_t n_draws,
                  const uint32_t* u, uint32_t* hap, uint32_t* entry) {
  return guarded([&] {
    require(fl && params && (n_draws == 0 || (u && hap)), "bad arguments");
    validate(*params);
    pcs_seq_params P = *params;
    P.shard_rank = 0;
    P.shard_count = 1;
    const HostPlan pl = make_host_plan(fl->host, P);
    for (const auto* v : {&pl.tiles, &pl.tiles_global})
      for (const pcs::Tile& t : *v) {
        if (t.id != tile_id) continue;
        const pcs::Entry* ent = pl.entries.data() + t.entry_off;
        const uint32_t* lo = pl.entry_lo.data() + t.entry_off;
        for (uint64_t i = 0; i < n_draws; ++i) {  // the device's place() / staged_read(), word for word
          uint32_t e = 0, base = 0;
          while (u[i] > ent[e].thr) {
            base = ent[e].thr + 1u;
            ++e;
          }
          hap[i] = fl->host.hap_list[ent[e].list_off + pcs::exact_leaf(u[i] - base, ent[e].scale, lo[e])];
          if (entry) entry[i] = e;
        }
        return;
      }
    throw std::domain_error("no tile with that id holds templates");
  });
}

}  // extern "C"
