// halo.cu -- halo slabs of the 2-D (x, y) domain decomposition.
//
// One launch moves the same (i, j, k) box of up to TB200_HALO_MAX_FIELDS fields between their
// storages and one contiguous message buffer, laid out [field][k][j][i] (i fastest), so that
// a side of the exchange costs one pack kernel, one NCCL send/recv pair (or one peer copy) and
// one unpack kernel regardless of the number of prognostic fields.  Pure data movement: the
// x-faces are 4-column slabs (32-byte runs per row), the y-faces full rows.
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
