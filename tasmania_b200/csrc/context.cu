// context.cu -- the scratch memory of the fused kernels behind an explicit handle (SURVEY.md section
// 8b: "scratch ... is owned by an explicit tb200_ctx_create / destroy handle held by the Python
// backend object").  The library never owns FIELD memory (that stays with the caller's allocator);
// what it needs beyond the fields -- the hand-off arrays of the fused RK stage (s after its first
// relaxation, the new Montgomery potential, the parked pressures of the column-scan variant) -- is
// requested from a context, which allocates it once per shape in the b200 storage layout (i fastest,
// rows padded to 16 doubles: the layout every kernel of the stage asserts) and frees it on destroy.
#include <map>
#include <vector>

#include "common.cuh"

using namespace tb200;

struct tb200_ctx {
  struct Key {
    int64_t n0, n1, n2;
    bool operator<(const Key &o) const {
      return n0 != o.n0 ? n0 < o.n0 : n1 != o.n1 ? n1 < o.n1 : n2 < o.n2;
    }
  };
  std::map<Key, std::vector<tb200_field>> scratch;  // per storage shape, in allocation order
  std::vector<void *> blocks;
};

extern "C" int tb200_ctx_create(tb200_ctx **ctx) {
  TB200_REQUIRE(ctx != nullptr, "ctx_create: NULL argument");
  *ctx = new tb200_ctx();
  return TB200_OK;
}

extern "C" int tb200_ctx_destroy(tb200_ctx *ctx) {
  if (ctx == nullptr) return TB200_OK;
  int rc = TB200_OK;
  for (void *p : ctx->blocks) {
    if (cudaFree(p) != cudaSuccess) {
      set_error("ctx_destroy: %s", cudaGetErrorString(cudaGetLastError()));
      rc = TB200_ERR_CUDA;
    }
  }
  delete ctx;
  return rc;
}

extern "C" int tb200_ctx_scratch(tb200_ctx *ctx, const int64_t shape[3], int count, tb200_field *fields) {
  TB200_REQUIRE(ctx != nullptr && shape != nullptr && fields != nullptr, "ctx_scratch: NULL argument");
  TB200_REQUIRE(count >= 1 && count <= 16 && shape[0] >= 1 && shape[1] >= 1 && shape[2] >= 1,
                "ctx_scratch: 1..16 fields of a non-empty shape");
  std::vector<tb200_field> &have = ctx->scratch[tb200_ctx::Key{shape[0], shape[1], shape[2]}];
  const int64_t pitch = (shape[0] + 15) / 16 * 16, plane = pitch * shape[1];
  while ((int)have.size() < count) {
    void *p = nullptr;
    const size_t bytes = (size_t)plane * (size_t)shape[2] * sizeof(double);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) {
      set_error("ctx_scratch(%lld x %lld x %lld): %s", (long long)shape[0], (long long)shape[1],
                (long long)shape[2], cudaGetErrorString(e));
      if (p != nullptr) cudaFree(p);
      return TB200_ERR_CUDA;
    }
    ctx->blocks.push_back(p);
    tb200_field f;
    f.ptr = p;
    f.shape[0] = shape[0]; f.shape[1] = shape[1]; f.shape[2] = shape[2];
    f.stride[0] = 1; f.stride[1] = pitch; f.stride[2] = plane;
    have.push_back(f);
  }
  for (int n = 0; n < count; ++n) fields[n] = have[n];
  return TB200_OK;
}
