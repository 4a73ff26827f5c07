// Byte-digit fixed-point representation shared by the exact int8 tensor-core contractions (oz8.cu) and the Gibbs tile
// kernels that emit K(X,Z) directly in it (gibbs_digits.cu).
//
// A real matrix X (R x Kd, rows padded to 128 or 64) with one power-of-two scale 2^e per row (or one for the whole matrix)
// is stored as the 56-bit two's-complement integer y = rint(x 2^(55-e)), |y| < 2^55, cut into NS = 7 bytes
//     x 2^-e = sum_p a_p 2^(-7-8p),   a_0 signed in [-128, 127] (top byte), a_1..a_6 unsigned in [0, 255],
// i.e. the digits ARE the bytes of y: no digit arithmetic, and 7 digits carry the same 56 bits as the 8 balanced base-128
// digits of round 1 -- 28 instead of 36 digit products with p + q <= 6.  tcgen05.mma kind::i8 takes the signedness of each
// operand from the instruction descriptor, so the top digit runs as S8 and the others as U8.
// Exactness of the int32 accumulators: |G_t| <= 7 * Kd * 255^2 < 2^31 for contraction lengths Kd <= 4608.
//
// Layout of the digit planes ("row layout", BR = rows per block: 128 for an A operand, 64 for a B operand), identical to the
// canonical K-major no-swizzle core-matrix order of the UMMA shared-memory descriptor, so that one pipeline stage (32
// contraction columns of all 7 digits of a row block) is ONE contiguous range of global memory:
//   offset(r, k, p) = ((((r / BR) * (Kd/32) + k/32) * 7 + p) * 2 + (k%32)/16) * (BR*16) + ((r%BR)/8) * 128 + (r%8) * 16 + k%16
// A core matrix is 8 rows x 16 consecutive columns = 128 contiguous bytes.  Read with the contraction along the columns
// (K-major) it feeds T = K C; read with the contraction along the ROWS (MN-major: the same core matrices, instruction
// descriptor a_major = b_major = 1) it feeds K^T K -- the SYRK needs no transposed copy when the scale is matrix-wide.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace npgp {

constexpr int O8_NS = 7;                            // digits (bytes) per entry
constexpr int O8_BM = 128;                          // rows of an A-operand block (TMEM lanes)
constexpr int O8_BN = 64;                           // rows of a B-operand block (TMEM columns per accumulator)
constexpr int O8_KS = 32;                           // contraction bytes per pipeline stage = one MMA k-step
constexpr int O8_A_STAGE = O8_NS * O8_BM * O8_KS;   // 28672 bytes
constexpr int O8_B_STAGE = O8_NS * O8_BN * O8_KS;   // 14336 bytes
constexpr int O8_POISON = 1 << 20;                  // exponent marking a row / column that holds a non-finite entry
constexpr int O8_EMIN = -960;                       // exponents are clamped here so that 2^(55-e) stays a normal double
constexpr int O8_MAX_KD = 4608;                     // longest exact contraction per int32 accumulation

__host__ __device__ inline long o8_digits_bytes(long rows, long Kd, int BR) {
  const long rpad = (rows + BR - 1) / BR * BR;
  return rpad * Kd * O8_NS;
}

// 2^k as a double (k in the normal range)
__device__ __forceinline__ double o8_pow2(int k) { return __hiloint2double((1023 + k) << 20, 0); }

// exponent e with |x| 2^-e < 1 for the largest magnitude `bits` (bit pattern of |x|, which orders like the value)
__device__ __forceinline__ int o8_exponent_of_max(unsigned long long bits) {
  if (bits >= 0x7FF0000000000000ull) return O8_POISON;  // inf / nan somewhere in the row
  const double m = __longlong_as_double((long long)bits);
  int e = 0;
  if (m > 0.0) frexp(m, &e);
  return e < O8_EMIN ? O8_EMIN : e;
}

// exponent of a positive scale s (device scalar), one more than frexp's when s sits at the top of its binade so that
// entries up to s (1 + 2^-40) still satisfy |x| 2^-e < 1
__device__ __forceinline__ int o8_exponent_of_scale(double s) {
  if (!(s > 0.0) || !(s < 1.7e308)) return O8_POISON;
  int e = 0;
  frexp(s * 1.0000000001, &e);
  return e < O8_EMIN ? O8_EMIN : e;
}

// Digit vectors of 16 consecutive contraction entries y[j] (|y| < 2^55): digit p of the 16 entries is one 16-byte vector,
// stored at base + p * plane_stride.  Digit p is byte (6 - p) of y; byte permutes only.
__device__ __forceinline__ void o8_store_digits(const long long (&y)[16], int8_t* base, long plane_stride) {
  uint32_t lo[16], hi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    lo[j] = (uint32_t)(unsigned long long)y[j];
    hi[j] = (uint32_t)((unsigned long long)y[j] >> 32);
  }
#pragma unroll
  for (int p = 0; p < O8_NS; ++p) {
    const int bi = (p < 3) ? (2 - p) : (6 - p);           // byte index inside hi (p < 3) or lo
    const uint32_t sel = (uint32_t)bi | ((uint32_t)(4 + bi) << 4);
    uint32_t w[4];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      const uint32_t* src = (p < 3) ? hi : lo;
      const uint32_t t0 = __byte_perm(src[4 * g4], src[4 * g4 + 1], sel);
      const uint32_t t1 = __byte_perm(src[4 * g4 + 2], src[4 * g4 + 3], sel);
      w[g4] = __byte_perm(t0, t1, 0x5410);
    }
    *reinterpret_cast<uint4*>(base + (long)p * plane_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// The inverse: entry c (0..3) of the 4 consecutive columns held in word w[p] of each digit plane -> y as a double
__device__ __forceinline__ double o8_digits_to_double(const uint32_t (&w)[O8_NS], int c) {
  const uint32_t sel = (uint32_t)c | ((uint32_t)(4 + c) << 4);
  const uint32_t t0 = __byte_perm(w[6], w[5], sel);   // [b6, b5, ., .]
  const uint32_t t1 = __byte_perm(w[4], w[3], sel);   // [b4, b3, ., .]
  const uint32_t lo = __byte_perm(t0, t1, 0x5410);
  const uint32_t t2 = __byte_perm(w[2], w[1], sel);   // [b2, b1, ., .]
  // byte 2 = b0, byte 3 = sign replication of b0 (selector nibble with bit 3 set)
  const uint32_t hi = __byte_perm(t2, w[0], 0x0010u | ((uint32_t)(4 + c) << 8) | ((uint32_t)(8 + 4 + c) << 12));
  return __ll2double_rn((long long)(((unsigned long long)hi << 32) | lo));
}

// y = rint(v * sc) limited to the 56-bit range (sc = 2^(55-e)); a non-finite product becomes 0 (the row is poisoned through
// its exponent instead)
__device__ __forceinline__ long long o8_quantise(double v, double sc) {
  long long y = __double2ll_rn(v * sc);
  const long long lim = (1ll << 55) - 1;
  y = y > lim ? lim : y;
  y = y < -lim ? -lim : y;
  return y;
}

}  // namespace npgp
