// A9: the secret-bit path as integer kernels.  Restates utils.py:74-97 (message <-> tensor)
// and the BER of train.py:285 of the reference, which run as float torch ops on CPU tensors.
// Messages are bit-packed (32 per uint32 word, LSB first); decode applies the reference's
// fp32 sequence  clamp -> +1 -> /step -> (>= 2^(sigma-i-1)) -> -=  verbatim, because the
// result is NOT `z >= 0`; errors are counted with XOR + popcount.
// __fmul_rn/__fadd_rn keep the compiler from contracting the sequence into FMAs.
#include "common.cuh"

namespace ideas {

__device__ __forceinline__ unsigned get_bit(const uint32_t* row, int j) { return (row[j >> 5] >> (j & 31)) & 1u; }

__global__ void __launch_bounds__(256) bits_encode_kernel(float* __restrict__ z, const uint32_t* __restrict__ bits,
                                                          const float* __restrict__ u, int B, int L, int sigma,
                                                          float step, float r, int words) {
  const int64_t total = (int64_t)B * L;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / L), l = (int)(idx - (int64_t)b * L);
    const uint32_t* row = bits + (int64_t)b * words;
    float nums = 0.f;                                        // utils.py:79-80, MSB first
    for (int i = 0; i < sigma; ++i) nums = __fadd_rn(nums, __fmul_rn((float)get_bit(row, l * sigma + i), (float)(1 << (sigma - i - 1))));
    float v = __fadd_rn(__fmul_rn(step, __fadd_rn(nums, 0.5f)), -1.f);   // utils.py:81
    const float uu = u ? u[idx] : 0.5f;
    const float jitter = __fadd_rn(__fmul_rn(__fmul_rn(uu, r), 2.f), -r);  // utils.py:82
    z[idx] = __fadd_rn(v, jitter);
  }
}

// one thread per 32-bit output word
__global__ void __launch_bounds__(256) bits_decode_kernel(uint32_t* __restrict__ bits, const float* __restrict__ z, int B,
                                                          int L, int sigma, float inv_step, int words) {
  const int64_t total = (int64_t)B * words;
  const int nbits = L * sigma;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / words), w = (int)(idx - (int64_t)b * words);
    uint32_t word = 0;
    int j = w * 32;
    const int jend = min(j + 32, nbits);
    while (j < jend) {
      const int l = j / sigma;
      float t = fminf(fmaxf(z[(int64_t)b * L + l], -1.f), 1.f);           // utils.py:89 clamp
      t = __fmul_rn(__fadd_rn(t, 1.f), inv_step);                          // +1, /step (step is a power of two)
      for (int i = 0; i < sigma; ++i) {                                    // utils.py:93-96
        const float wgt = (float)(1 << (sigma - i - 1));
        const unsigned bit = t >= wgt ? 1u : 0u;
        t = __fadd_rn(t, -__fmul_rn((float)bit, wgt));
        const int jj = l * sigma + i;
        if (jj >= w * 32 && jj < jend) word |= bit << (jj & 31);
      }
      j = (l + 1) * sigma;
    }
    bits[idx] = word;
  }
}

__global__ void __launch_bounds__(256) bits_count_errors_kernel(unsigned long long* __restrict__ errors,
                                                                const uint32_t* __restrict__ a,
                                                                const uint32_t* __restrict__ b, int64_t n) {
  unsigned long long local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    local += __popc(a[i] ^ b[i]);
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(errors, local);
}

}  // namespace ideas

using namespace ideas;

static int bits_blocks(int64_t items) {
  int64_t b = ceil_div64(items, 256);
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return b < 1 ? 1 : (int)b;
}

extern "C" int ideas_bits_encode(float* z, const uint32_t* bits, const float* u, int B, int L, int sigma, float delta,
                                 void* stream) {
  IDEAS_REQUIRE(B >= 0 && L >= 0 && sigma >= 1 && sigma <= 16, "bits_encode: bad shape (sigma must be 1..16)");
  if ((int64_t)B * L == 0) return IDEAS_OK;
  IDEAS_REQUIRE(z && bits, "bits_encode: null pointer");
  const float step = 2.0f / (float)(1 << sigma);
  const float r = (float)((double)step * (double)delta);   // python float product, then fp32 tensor op
  const int words = (L * sigma + 31) / 32;
  bits_encode_kernel<<<bits_blocks((int64_t)B * L), 256, 0, (cudaStream_t)stream>>>(z, bits, u, B, L, sigma, step, r, words);
  IDEAS_CHECK_LAUNCH("bits_encode");
  return IDEAS_OK;
}

extern "C" int ideas_bits_decode(uint32_t* bits, const float* z, int B, int L, int sigma, void* stream) {
  IDEAS_REQUIRE(B >= 0 && L >= 0 && sigma >= 1 && sigma <= 16, "bits_decode: bad shape (sigma must be 1..16)");
  if ((int64_t)B * L == 0) return IDEAS_OK;
  IDEAS_REQUIRE(z && bits, "bits_decode: null pointer");
  const float inv_step = (float)(1 << sigma) / 2.0f;
  const int words = (L * sigma + 31) / 32;
  bits_decode_kernel<<<bits_blocks((int64_t)B * words), 256, 0, (cudaStream_t)stream>>>(bits, z, B, L, sigma, inv_step, words);
  IDEAS_CHECK_LAUNCH("bits_decode");
  return IDEAS_OK;
}

extern "C" int ideas_bits_count_errors(unsigned long long* errors, const uint32_t* a, const uint32_t* b, int64_t n_words,
                                       void* stream) {
  IDEAS_REQUIRE(n_words >= 0, "bits_count_errors: negative size");
  if (n_words == 0) return IDEAS_OK;
  IDEAS_REQUIRE(errors && a && b, "bits_count_errors: null pointer");
  bits_count_errors_kernel<<<bits_blocks(n_words), 256, 0, (cudaStream_t)stream>>>(errors, a, b, n_words);
  IDEAS_CHECK_LAUNCH("bits_count_errors");
  return IDEAS_OK;
}
