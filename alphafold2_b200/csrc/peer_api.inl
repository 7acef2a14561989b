// =================================================================================================
// Peer-memory exchange for the sharded trunk (alphafold2_b200/parallel.py): the row-shard <-> column-shard
// re-layout of the pair and MSA tensors (the all-to-all of the FastFold-style schedule) done by ONE kernel per
// exchange that stores each chunk straight into its destination rank's buffer over NVLink and then joins a
// flag barrier, instead of pack kernel + NCCL all_to_all + unpack kernel.
//
// Every rank owns one `cudaMalloc`ed arena that the other ranks of the node map through CUDA IPC:
//   [0, 4096)   control block: u32 flags[2][32] (arrival epochs, written by peers), u32 epoch[2] at +1024,
//               u32 done[2] at +1088 (CTA completion counters), u32 err at +1152
//   [4096, ...) the row-layout and column-layout buffers of the two tracks (pair, MSA)
// Included at the end of api.cu (shares its helpers).
// =================================================================================================
namespace af2 {

constexpr int PEER_CTRL_BYTES = 4096;
constexpr int PEER_VEC = 4;              // 16-byte vectors per thread: a CTA moves 256 x PEER_VEC x 16 B of one row
constexpr int PEER_MAX_RANKS = 32;

struct PeerExchangeParams {
  const char* src;              // local source
  long long src_peer_stride;    // chunk for rank p starts at src + p * src_peer_stride
  long long src_row_stride;     // bytes between consecutive rows of a chunk
  char* const* peer_base;       // device array [P]: arena base of every rank as mapped here (own arena included)
  long long dst_off;            // where this rank's chunk starts inside the destination arena
  long long dst_row_stride;
  long long row_bytes;          // contiguous bytes per row (multiple of 16)
  int rank, P, channel, interleave;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// grid (segments x P, rows); block 256.  A CTA moves one segment of one row of the chunk bound for one rank; the
// destination is rotated by the sender's rank so that the P senders address P different receivers at any moment.  The CTA that finishes last signals
// "my stores are done" to every rank (release, system scope) and waits for the same signal from every rank, so when the
// kernel completes every chunk bound for THIS rank has landed and the stream's next kernel may read the buffer.
template <int V, bool FENCE_ALL>
__global__ void __launch_bounds__(256) peer_exchange_kernel(const PeerExchangeParams p) {
  // interleave: the destination rank is the fastest-varying CTA index, so the CTAs resident at any moment address all
  // P ranks (the local copy overlaps the NVLink stores); otherwise the grid walks the destinations one after another
  const int peer = p.interleave ? (int)((blockIdx.x % p.P + p.rank) % p.P) : (int)((blockIdx.z + p.rank) % p.P);
  const unsigned seg_i = p.interleave ? blockIdx.x / p.P : blockIdx.x;
  const long long off0 = (long long)seg_i * (V * 4096) + threadIdx.x * 16;
  const char* s = p.src + (long long)peer * p.src_peer_stride + (long long)blockIdx.y * p.src_row_stride;
  char* d = p.peer_base[peer] + p.dst_off + (long long)blockIdx.y * p.dst_row_stride;
  uint4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const long long o = off0 + i * 4096;
    if (o < p.row_bytes) v[i] = *reinterpret_cast<const uint4*>(s + o);
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const long long o = off0 + i * 4096;
    if (o < p.row_bytes) *reinterpret_cast<uint4*>(d + o) = v[i];
  }
  // FENCE_ALL: every thread makes its own stores visible system-wide before the CTA counts itself done.  Otherwise the
  // CTA barrier orders the stores before thread 0, whose system-scope fence is cumulative over them (PTX memory model:
  // bar.sync synchronises the CTA's threads, the fence then covers every write that happened before it in causality order).
  if (FENCE_ALL) __threadfence_system();
  __syncthreads();
  __shared__ int is_last;
  char* mine = p.peer_base[p.rank];
  unsigned* done = reinterpret_cast<unsigned*>(mine + 1088) + p.channel;
  if (threadIdx.x == 0) {
    if (!FENCE_ALL) __threadfence_system();
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    is_last = atomicAdd(done, 1u) == total - 1;
  }
  __syncthreads();
  if (!is_last || threadIdx.x >= 32) return;
  __threadfence();
  unsigned* epoch = reinterpret_cast<unsigned*>(mine + 1024) + p.channel;
  const unsigned e = *epoch + 1;
  __syncwarp();
  if ((int)threadIdx.x < p.P) {
    unsigned* theirs = reinterpret_cast<unsigned*>(p.peer_base[threadIdx.x]) + p.channel * PEER_MAX_RANKS + p.rank;
    st_release_sys(theirs, e);
    const unsigned* slot = reinterpret_cast<const unsigned*>(mine) + p.channel * PEER_MAX_RANKS + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(slot) - e) < 0) {
      if (global_ns() - t0 > 20ull * 1000 * 1000 * 1000) {          // a rank never arrived: give up instead of hanging the GPU
        *reinterpret_cast<unsigned*>(mine + 1152) = 1u;
        break;
      }
    }
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    *epoch = e;
    *done = 0u;
  }
}

}  // namespace af2

extern "C" {

int af2_peer_ctrl_bytes(void) { return af2::PEER_CTRL_BYTES; }

int af2_peer_can_access(int device, int peer_device) {
  int ok = 0;
  if (device == peer_device) return 1;
  if (cudaDeviceCanAccessPeer(&ok, device, peer_device) != cudaSuccess) { cudaGetLastError(); return 0; }
  return ok;
}

int af2_peer_alloc(long long bytes, void** ptr) {
  using namespace af2;
  if (!ptr || bytes < PEER_CTRL_BYTES) return fail(AF2_ERR_BAD_ARG, "peer_alloc: bad argument");
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return fail(AF2_ERR_CUDA, "peer_alloc: %s", cudaGetErrorString(e)); }
  *ptr = p;
  return AF2_OK;
}

int af2_peer_free(void* ptr) {
  using namespace af2;
  if (ptr) CUDA_OK(cudaFree(ptr));
  return AF2_OK;
}

int af2_peer_export(const void* ptr, unsigned char* handle64) {
  using namespace af2;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle64) return fail(AF2_ERR_BAD_ARG, "peer_export: bad argument");
  cudaIpcMemHandle_t h;
  CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, 64);
  return AF2_OK;
}

int af2_peer_open(const unsigned char* handle64, void** ptr) {
  using namespace af2;
  if (!ptr || !handle64) return fail(AF2_ERR_BAD_ARG, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return AF2_OK;
}

int af2_peer_close(void* ptr) {
  using namespace af2;
  if (ptr) CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return AF2_OK;
}

// 1 if a barrier of this arena ever timed out (a rank did not arrive within 20 s); synchronises the device
int af2_peer_error(const void* my_base) {
  unsigned v = 0;
  if (!my_base) return 0;
  if (cudaMemcpy(&v, static_cast<const char*>(my_base) + 1152, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 1; }
  return (int)v;
}

int af2_peer_exchange(const void* src, long long src_peer_stride, long long src_row_stride, void* const* peer_base,
                      long long dst_off, long long dst_row_stride, int rows, long long row_bytes, int channel, int rank, int P,
                      af2_stream_t stream) {
  using namespace af2;
  NvtxRange nvtx_("af2_peer_exchange");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!src || !peer_base || rows < 1 || row_bytes < 16 || P < 1 || P > PEER_MAX_RANKS || rank < 0 || rank >= P || channel < 0 || channel > 1)
    return fail(AF2_ERR_BAD_ARG, "peer_exchange: bad argument");
  if ((row_bytes | src_peer_stride | src_row_stride | dst_off | dst_row_stride | (long long)reinterpret_cast<uintptr_t>(src)) & 15)
    return fail(AF2_ERR_BAD_ARG, "peer_exchange: rows, strides and offsets must be multiples of 16 bytes");
  if (rows > 65535) return fail(AF2_ERR_BAD_ARG, "peer_exchange: %d rows per chunk (max 65535)", rows);
  PeerExchangeParams p;
  p.src = static_cast<const char*>(src); p.src_peer_stride = src_peer_stride; p.src_row_stride = src_row_stride;
  p.peer_base = reinterpret_cast<char* const*>(peer_base); p.dst_off = dst_off; p.dst_row_stride = dst_row_stride;
  p.row_bytes = row_bytes; p.rank = rank; p.P = P; p.channel = channel;
  const char* ev = getenv("AF2_PEER_VARIANT");   // experiment knob (tools/peer_bench.py): bit 0 = 8 vectors per thread, bit 1 = one fence per CTA, bit 2 = destinations one after another
  const int variant = ev ? atoi(ev) : 2;    // default: 4 vectors per thread, one fence per CTA, destinations interleaved (profiles/r02p_peer_bench_2gpu.log)
  const int vec = (variant & 1) ? 8 : PEER_VEC;
  const long long seg = (long long)vec * 4096;
  p.interleave = (variant & 4) ? 0 : 1;
  const unsigned nseg = (unsigned)((row_bytes + seg - 1) / seg);
  dim3 grid(p.interleave ? nseg * P : nseg, (unsigned)rows, p.interleave ? 1u : (unsigned)P);
  const double bytes = (double)rows * (double)row_bytes * P;
  ProfScope ps(s, 5, 0.0, 2.0 * bytes);
  switch (variant & 3) {
    case 1: peer_exchange_kernel<8, true><<<grid, 256, 0, s>>>(p); break;
    case 2: peer_exchange_kernel<PEER_VEC, false><<<grid, 256, 0, s>>>(p); break;
    case 3: peer_exchange_kernel<8, false><<<grid, 256, 0, s>>>(p); break;
    default: peer_exchange_kernel<PEER_VEC, true><<<grid, 256, 0, s>>>(p); break;
  }
  CUDA_OK(cudaGetLastError());
  return AF2_OK;
}

}  // extern "C"
