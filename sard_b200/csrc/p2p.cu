// Gradient all-reduce of the data-parallel update over NVLink / NVSwitch peer memory, in ONE kernel on the net's own
// stream (the transition-sharded PPO minibatch ends in a SUM of the ranks' gradient buckets).
//
// Two-shot, in place, deterministic: rank r owns the r-th slice of the bucket.  After a barrier ("every rank's backward
// has finished") it reads that slice from every rank's bucket with direct peer loads, adds the copies in rank order
// (so every element is reduced exactly once, by one rank, in a fixed order: all ranks end up with bit-identical sums and
// reruns are bit-identical) and writes the sum into every rank's bucket with direct peer stores.  A second barrier
// ("every slice has been written everywhere") ends the kernel; the optimiser kernels that follow read local memory.
// Against the NCCL path this removes the two cross-stream hops into and out of the communicator's stream and the
// collective's own launch chain: ~15 us instead of ~60 us per minibatch and net on 8 GPUs.
//
// Barriers: one 32-bit slot per (block, peer) in the symmetric signal pad, written with st.release.sys and polled with
// ld.acquire.sys; the expected value increases monotonically with every call (no reset, no ABA).  Block b only ever
// talks to block b of the peers.  The polls are bounded: a missing peer traps the kernel instead of hanging the GPU.
#include "common.cuh"

#define P2P_MAX_WORLD 8
#define P2P_BLOCKS 24
#define P2P_THREADS 512

struct P2pParams {
  float* const* bufs;            // device array [world]: every rank's bucket (peer-mapped)
  uint32_t* const* pads;         // device array [world]: every rank's signal pad
  int rank, world;
  long long n4;                  // float4 elements to reduce
  uint32_t v0;                   // value of the first barrier of this call (the second uses v0 + 1)
};

__device__ __forceinline__ void p2p_put(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t p2p_get(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_barrier(const P2pParams& p, uint32_t v) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    p2p_put(p.pads[peer] + blockIdx.x * P2P_MAX_WORLD + p.rank, v);
    const uint32_t* mine = p.pads[p.rank] + blockIdx.x * P2P_MAX_WORLD + peer;
    long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
      if ((int)(p2p_get(mine) - v) >= 0) break;
      if ((spin & 0xfff) == 0xfff) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 6000000000LL) {              // ~3 s
          printf("sard_b200: p2p all-reduce barrier timed out (rank %d block %d peer %d want %u have %u)\n", p.rank, blockIdx.x, peer, v,
                 p2p_get(mine));
          __trap();
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(P2P_THREADS) allreduce_p2p_kernel(const P2pParams p) {
  pdl_sync();
  p2p_barrier(p, p.v0);                                   // every rank's bucket is complete
  const long long chunk = (p.n4 + p.world - 1) / p.world;
  const long long lo = chunk * p.rank, hi = lo + chunk < p.n4 ? lo + chunk : p.n4;
  float4* bufs[P2P_MAX_WORLD];
#pragma unroll
  for (int r = 0; r < P2P_MAX_WORLD; ++r) bufs[r] = reinterpret_cast<float4*>(p.bufs[r < p.world ? r : 0]);
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
    float4 v[P2P_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < P2P_MAX_WORLD; ++r)
      if (r < p.world) v[r] = bufs[r][i];                 // all peer loads in flight together
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < P2P_MAX_WORLD; ++r)
      if (r < p.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }      // rank order
#pragma unroll
    for (int r = 0; r < P2P_MAX_WORLD; ++r)
      if (r < p.world) bufs[r][i] = s;
  }
  p2p_barrier(p, p.v0 + 1);                               // every slice has landed in every bucket
}

extern "C" int sard_allreduce_p2p_pad_bytes(void) { return P2P_BLOCKS * P2P_MAX_WORLD * 4; }

extern "C" int sard_allreduce_p2p(const void* bufs_dev, const void* pads_dev, int rank, int world, long long n_floats,
                                  unsigned first_barrier_value, sard_stream_t stream) {
  SARD_REQUIRE(world >= 2 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "p2p all-reduce: world size 2..8");
  SARD_REQUIRE(n_floats > 0 && n_floats % 4 == 0, "p2p all-reduce: length must be a multiple of 4 floats");
  P2pParams p;
  p.bufs = reinterpret_cast<float* const*>(bufs_dev); p.pads = reinterpret_cast<uint32_t* const*>(pads_dev);
  p.rank = rank; p.world = world; p.n4 = n_floats / 4; p.v0 = first_barrier_value;
  sard_prof_bytes(4.0 * n_floats * 2.0 * (world - 1) / world);       // peer reads + peer writes of one rank
  sard_launch(allreduce_p2p_kernel, P2P_BLOCKS, P2P_THREADS, 0, as_stream(stream), p);
  SARD_LAUNCH_OK("allreduce_p2p_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------ moments all-gather
// All-gather of one small record (the RunningNorm moments of a run: 1 + 2D doubles) over peer memory, for the prefetched
// gather phase of the data-parallel update.  Every rank owns a symmetric "board"; record r of the (slot, key) chunk at
// `off` is written by rank r into EVERY rank's board with direct peer stores.  Barrier 1: every rank has reached this
// all-gather, i.e. the previous reader of the chunk on that rank has finished (its gather stream waited for the
// workspace's `free` event first); barrier 2: every record has landed everywhere.  One block; replaces an NCCL
// all-gather issued through torch.distributed, whose HOST cost (~90 us per call, two calls per minibatch) made the
// data-parallel update host-bound (117 ms of enqueue time for a 123 ms update on 2 GPUs).
struct AgParams {
  double* const* boards;         // device array [world]: every rank's board (peer-mapped)
  uint32_t* const* pads;         // device array [world]: every rank's signal pad (its own allocation's)
  int rank, world, rec;
  long long off;                 // doubles from the board base to the chunk [world][rec]
  const double* local;           // this rank's record
  uint32_t v0;
};
__global__ void __launch_bounds__(256) allgather_p2p_kernel(const AgParams a) {
  pdl_sync();
  P2pParams b;
  b.bufs = nullptr; b.pads = a.pads; b.rank = a.rank; b.world = a.world; b.n4 = 0; b.v0 = a.v0;
  p2p_barrier(b, a.v0);
  for (int peer = 0; peer < a.world; ++peer) {
    double* dst = a.boards[peer] + a.off + (long long)a.rank * a.rec;
    for (int i = threadIdx.x; i < a.rec; i += blockDim.x) dst[i] = a.local[i];
  }
  p2p_barrier(b, a.v0 + 1);
}
extern "C" int sard_allgather_p2p_pad_bytes(void) { return P2P_MAX_WORLD * 4; }
extern "C" int sard_allgather_p2p(const void* boards_dev, const void* pads_dev, int rank, int world, long long chunk_offset_doubles,
                                  const double* local_record, int record_doubles, unsigned first_barrier_value,
                                  sard_stream_t stream) {
  SARD_REQUIRE(world >= 2 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "p2p all-gather: world size 2..8");
  SARD_REQUIRE(record_doubles > 0 && chunk_offset_doubles >= 0 && local_record != nullptr, "p2p all-gather: record");
  AgParams a;
  a.boards = reinterpret_cast<double* const*>(boards_dev); a.pads = reinterpret_cast<uint32_t* const*>(pads_dev);
  a.rank = rank; a.world = world; a.rec = record_doubles; a.off = chunk_offset_doubles; a.local = local_record;
  a.v0 = first_barrier_value;
  sard_launch(allgather_p2p_kernel, 1, 256, 0, as_stream(stream), a);
  SARD_LAUNCH_OK("allgather_p2p_kernel");
  return 0;
}
