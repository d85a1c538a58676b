// halo.cu -- halo slabs of the 2-D (x, y) domain decomposition.
//
// One launch moves the same (i, j, k) box of up to TB200_HALO_MAX_FIELDS fields between their
// storages and one contiguous message buffer, laid out [field][k][j][i] (i fastest), so that
// a side of the exchange costs one pack kernel, one NCCL send/recv pair (or one peer copy) and
// one unpack kernel regardless of the number of prognostic fields.  Pure data movement: the
// x-faces are 4-column slabs (32-byte runs per row), the y-faces full rows.
#include <string.h>

#include "common.cuh"

using namespace tb200;

namespace {

struct HaloFields {
  View f[TB200_HALO_MAX_FIELDS];
  int n;
};

template <bool PACK>
__global__ void __launch_bounds__(256) halo_kernel(const HaloFields hf, double *buffer, int i0,
                                                   int j0, int k0, int di, int dj, int dk) {
  // threadIdx.x runs along i when the box is wide, along j*i otherwise: a linear index over the
  // (i, j) plane keeps all 32 lanes busy on the narrow x-slabs
  const long long plane = (long long)di * dj;
  const long long box = plane * dk;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < plane;
       p += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(p / di), i = (int)(p - (long long)j * di);
    for (int k = blockIdx.y; k < dk; k += gridDim.y) {
      const long long b = p + plane * k;
#pragma unroll
      for (int n = 0; n < TB200_HALO_MAX_FIELDS; ++n) {
        if (n < hf.n) {
          if (PACK)
            buffer[b + box * n] = hf.f[n](i + i0, j + j0, k + k0);
          else
            hf.f[n](i + i0, j + j0, k + k0) = buffer[b + box * n];
        }
      }
    }
  }
}

template <bool PACK>
int run_halo(const tb200_field *const *fields, int nfields, double *buffer,
             const int32_t o[3], const int32_t d[3], void *stream, const char *what) {
  TB200_REQUIRE(fields != nullptr && buffer != nullptr, "%s: NULL argument", what);
  TB200_REQUIRE(nfields >= 1 && nfields <= TB200_HALO_MAX_FIELDS, "%s: 1..%d fields, got %d", what,
                TB200_HALO_MAX_FIELDS, nfields);
  HaloFields hf{};
  hf.n = nfields;
  for (int n = 0; n < nfields; ++n) {
    hf.f[n] = view(fields[n]);
    TB200_REQUIRE(box_inside(hf.f[n], o, d), "%s: box outside the storage of field %d", what, n);
  }
  if (d[0] == 0 || d[1] == 0 || d[2] == 0) return TB200_OK;
  const long long plane = (long long)d[0] * d[1];
  dim3 block(256, 1, 1);
  dim3 grid((unsigned)((plane + 255) / 256 > 4096 ? 4096 : (plane + 255) / 256),
            (unsigned)(d[2] > 65535 ? 65535 : d[2]), 1);
  halo_kernel<PACK><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      hf, buffer, o[0], o[1], o[2], d[0], d[1], d[2]);
  return check_launch(what);
}

}  // namespace

extern "C" int tb200_halo_pack(const tb200_field *const *fields, int nfields, double *buffer,
                               const int32_t origin[3], const int32_t domain[3], void *stream) {
  return run_halo<true>(fields, nfields, buffer, origin, domain, stream, "halo_pack");
}

extern "C" int tb200_halo_unpack(const tb200_field *const *fields, int nfields,
                                 const double *buffer, const int32_t origin[3],
                                 const int32_t domain[3], void *stream) {
  return run_halo<false>(fields, nfields, const_cast<double *>(buffer), origin, domain, stream,
                         "halo_unpack");
}

// ---------------------------------------------------------------------------------------------
// Peer-to-peer transport (one process per GPU, NVLink / NVSwitch): the pack kernel of a side
// stores the slab STRAIGHT into the neighbour's receive buffer through a CUDA-IPC mapping of the
// neighbour's memory and then raises the neighbour's arrival counter (system-scope release); the
// neighbour's unpack kernel waits for the counter (system-scope acquire) and scatters the slab
// into its halo.  No send buffer, no NCCL call, no host synchronisation: one `push` and one
// `pull` launch for ALL sides of an exchange (the four faces and the four corner blocks of a
// single-phase exchange, or the two faces of one phase of a two-phase one), capturable in a CUDA
// graph.
//
// Flow control without acknowledgements: a receive buffer has two slots, exchange number q goes
// to slot q & 1.  A rank can only push exchange q + 2 after it has pulled exchange q + 1 from the
// same neighbour, which that neighbour pushed after (stream order) it had finished pulling
// exchange q -- so slot q & 1 is free again by then.  The sequence numbers live in device memory
// (tb200_halo_channel) and are advanced by the kernels themselves, so a captured graph replays
// correctly.  A wait that does not complete within ~20 s sets channel.error (a crashed peer must
// not hang the box); tb200_p2p_channel_error reports it.
namespace {

struct Channel {              // device memory, one per (exchange object, phase)
  unsigned long long pushed;  // exchanges pushed so far
  unsigned long long pulled;  // exchanges pulled so far
  unsigned int done_push, done_pull;  // blocks that have finished the current launch
  unsigned int error, pad;
};
static_assert(sizeof(Channel) == TB200_P2P_CHANNEL_BYTES, "channel size");

struct SideDev {
  double *remote_buffer;               // neighbour's receive buffer for the slab I send (2 slots)
  unsigned long long *remote_counter;  // neighbour's arrival counter of that buffer
  double *local_buffer;                // my receive buffer of this side (2 slots)
  unsigned long long *local_counter;   // my arrival counter (raised by the neighbour)
  long long slot_doubles;
  int si0, sj0, ri0, rj0, di, dj;      // send / receive origin and extent in (i, j)
};
struct Sides {
  SideDev s[TB200_HALO_MAX_SIDES];
  int n;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid (blocks per side, sides); the threads of a side walk its box linearly, i fastest, then j,
// then k (message layout [field][k][j][i] as above), each moving its point of every field
template <bool PUSH>
__global__ void __launch_bounds__(256) p2p_kernel(const HaloFields hf, const Sides sides, Channel *ch,
                                                  int k0, int dk, unsigned long long spin_limit) {
  const SideDev sd = sides.s[blockIdx.y];
  const unsigned long long q = (PUSH ? ch->pushed : ch->pulled) + 1ull;  // this exchange
  double *buffer = (PUSH ? sd.remote_buffer : sd.local_buffer) + (long long)(q & 1ull) * sd.slot_doubles;
  if (!PUSH) {  // the slab of exchange q must have landed
    if (threadIdx.x == 0) {
      unsigned long long spins = 0;
      while (ld_acquire_sys(sd.local_counter) < q) {
        __nanosleep(200);
        if (++spins > spin_limit) {  // the peer is gone: flag it and carry on (no hang)
          atomicExch(&ch->error, 1u);
          break;
        }
      }
    }
    __syncthreads();
  }
  const int i0 = PUSH ? sd.si0 : sd.ri0, j0 = PUSH ? sd.sj0 : sd.rj0;
  const long long plane = (long long)sd.di * sd.dj;
  const long long box = plane * dk;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < box;
       b += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(b / plane);
    const long long p = b - (long long)k * plane;
    const int j = (int)(p / sd.di), i = (int)(p - (long long)j * sd.di);
#pragma unroll
    for (int n = 0; n < TB200_HALO_MAX_FIELDS; ++n) {
      if (n < hf.n) {
        if (PUSH)
          buffer[b + box * n] = hf.f[n](i + i0, j + j0, k + k0);
        else  // written by the peer: read around L1
          hf.f[n](i + i0, j + j0, k + k0) = __ldcg(buffer + (b + box * n));
      }
    }
  }
  // the last block of the launch publishes: a push raises the neighbours' arrival counters (after
  // every block's stores are visible system-wide), both advance the channel's sequence number
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int *done = PUSH ? &ch->done_push : &ch->done_pull;
    const unsigned int total = gridDim.x * gridDim.y;
    if (atomicAdd(done, 1u) == total - 1u) {
      __threadfence_system();
      if (PUSH) {
        for (int n = 0; n < sides.n; ++n) st_release_sys(sides.s[n].remote_counter, q);
        ch->pushed = q;
      } else {
        ch->pulled = q;
      }
      *done = 0u;
      __threadfence();
    }
  }
}

template <bool PUSH>
int run_p2p(const tb200_field *const *fields, int nfields, const tb200_halo_side *sides, int nsides,
            void *channel, int k0, int dk, void *stream, const char *what) {
  TB200_REQUIRE(fields != nullptr && sides != nullptr && channel != nullptr, "%s: NULL argument", what);
  TB200_REQUIRE(nfields >= 1 && nfields <= TB200_HALO_MAX_FIELDS, "%s: 1..%d fields, got %d", what,
                TB200_HALO_MAX_FIELDS, nfields);
  TB200_REQUIRE(nsides >= 0 && nsides <= TB200_HALO_MAX_SIDES, "%s: 0..%d sides per launch, got %d", what,
                TB200_HALO_MAX_SIDES, nsides);
  if (nsides == 0 || dk <= 0) return TB200_OK;
  HaloFields hf{};
  hf.n = nfields;
  Sides sd{};
  sd.n = nsides;
  long long plane_max = 0;
  for (int n = 0; n < nfields; ++n) hf.f[n] = view(fields[n]);
  for (int m = 0; m < nsides; ++m) {
    const tb200_halo_side &h = sides[m];
    TB200_REQUIRE(h.remote_buffer != nullptr && h.remote_counter != nullptr && h.local_buffer != nullptr &&
                      h.local_counter != nullptr, "%s: side %d has a NULL buffer / counter", what, m);
    TB200_REQUIRE(h.extent[0] > 0 && h.extent[1] > 0, "%s: side %d is empty", what, m);
    const int32_t o[3] = {PUSH ? h.send_origin[0] : h.recv_origin[0],
                          PUSH ? h.send_origin[1] : h.recv_origin[1], k0};
    const int32_t d[3] = {h.extent[0], h.extent[1], dk};
    for (int n = 0; n < nfields; ++n)
      TB200_REQUIRE(box_inside(hf.f[n], o, d), "%s: side %d: box outside the storage of field %d", what, m, n);
    TB200_REQUIRE((long long)nfields * d[0] * d[1] * dk <= h.slot_doubles,
                  "%s: side %d: message larger than a buffer slot", what, m);
    sd.s[m] = SideDev{h.remote_buffer, reinterpret_cast<unsigned long long *>(h.remote_counter),
                      h.local_buffer, reinterpret_cast<unsigned long long *>(h.local_counter),
                      h.slot_doubles, h.send_origin[0], h.send_origin[1], h.recv_origin[0],
                      h.recv_origin[1], h.extent[0], h.extent[1]};
    plane_max = plane_max > (long long)d[0] * d[1] ? plane_max : (long long)d[0] * d[1];
  }
  // enough blocks to fill the machine with both sides (a waiting block only parks one thread and
  // depends on the peer alone, never on another block of its own launch)
  const long long want = (plane_max * dk + 255) / 256;
  const unsigned bx = (unsigned)(want > 4 * 148 ? 4 * 148 : want);
  dim3 grid(bx, (unsigned)nsides, 1);
  // ~200 ns per poll: 1e8 polls = 20 s
  p2p_kernel<PUSH><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      hf, sd, static_cast<Channel *>(channel), k0, dk, 100000000ull);
  return check_launch(what);
}

}  // namespace

extern "C" int tb200_halo_push(const tb200_field *const *fields, int nfields,
                               const tb200_halo_side *sides, int nsides, void *channel, int k0,
                               int nk, void *stream) {
  return run_p2p<true>(fields, nfields, sides, nsides, channel, k0, nk, stream, "halo_push");
}

extern "C" int tb200_halo_pull(const tb200_field *const *fields, int nfields,
                               const tb200_halo_side *sides, int nsides, void *channel, int k0,
                               int nk, void *stream) {
  return run_p2p<false>(fields, nfields, sides, nsides, channel, k0, nk, stream, "halo_pull");
}

extern "C" int tb200_p2p_alloc(size_t bytes, void **ptr) {
  TB200_REQUIRE(ptr != nullptr && bytes > 0, "p2p_alloc: NULL / empty request");
  // plain cudaMalloc: exportable with cudaIpcGetMemHandle (a caching allocator's sub-block or a
  // virtual-memory segment is not)
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("p2p_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    return TB200_ERR_CUDA;
  }
  return TB200_OK;
}

extern "C" int tb200_p2p_free(void *ptr) {
  if (ptr != nullptr && cudaFree(ptr) != cudaSuccess) {
    set_error("p2p_free: %s", cudaGetErrorString(cudaGetLastError()));
    return TB200_ERR_CUDA;
  }
  return TB200_OK;
}

extern "C" int tb200_p2p_export(void *ptr, unsigned char handle[TB200_P2P_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == TB200_P2P_HANDLE_BYTES, "handle size");
  TB200_REQUIRE(ptr != nullptr && handle != nullptr, "p2p_export: NULL argument");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) {
    set_error("p2p_export: %s", cudaGetErrorString(cudaGetLastError()));
    return TB200_ERR_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  return TB200_OK;
}

extern "C" int tb200_p2p_import(const unsigned char handle[TB200_P2P_HANDLE_BYTES], void **ptr) {
  TB200_REQUIRE(ptr != nullptr && handle != nullptr, "p2p_import: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    set_error("p2p_import: %s (peer access over NVLink / PCIe is required between the two devices)",
              cudaGetErrorString(cudaGetLastError()));
    return TB200_ERR_CUDA;
  }
  return TB200_OK;
}

extern "C" int tb200_p2p_release(void *ptr) {
  if (ptr != nullptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) {
    set_error("p2p_release: %s", cudaGetErrorString(cudaGetLastError()));
    return TB200_ERR_CUDA;
  }
  return TB200_OK;
}

extern "C" int tb200_p2p_channel_error(const void *channel, int *error) {
  TB200_REQUIRE(channel != nullptr && error != nullptr, "p2p_channel_error: NULL argument");
  Channel c;
  if (cudaMemcpy(&c, channel, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("p2p_channel_error: %s", cudaGetErrorString(cudaGetLastError()));
    return TB200_ERR_CUDA;
  }
  *error = (int)c.error;
  return TB200_OK;
}
