// Byte-digit fixed-point representation shared by the exact int8 tensor-core contractions (oz8.cu) and the Gibbs tile
// kernels that emit K(X,Z) directly in it (gibbs_digits.cu).
//
// A real matrix X (R x Kd, rows padded to 128 or 64) with one power-of-two scale 2^e per row (or one for the whole matrix)
// is stored as the integer y = rint(x 2^(54-e)), |y| <= 2^54, written in BALANCED base 256:
//     x 2^-e = sum_{p=0..6} a_p 2^(-6-8p),      a_p in [-128, 127]   (all seven digits signed)
// (digit p is byte 6-p of y + 0x0000808080808080 with the low six bytes' top bit flipped: two integer operations per word).
// Seven signed digits carry 55 bits -- the same resolution relative to the row maximum as a double has relative to 1/4 of
// its own value -- and need 28 digit products with p + q <= 6 where round 1's eight 7-bit digits needed 36.  Every MMA is
// S8 x S8 with one constant instruction descriptor.
// Exactness of the int32 accumulators: |G_t| <= 7 * Kd * 128^2 < 2^31 for contraction lengths Kd <= 18724.
//
// Layout of the digit planes of an A-type operand (row blocks of 128; "row layout"):
//   offset(r, k, p) = (((r/128) * 7 + p) * (Kd/32) + k/32) * 4096 + ((k%32)/16) * 2048 + ((r%128)/8) * 128 + (r%8) * 16 + k%16
// i.e. per row block and digit a sequence of k-steps, each k-step two 2048-byte planes (the two 16-byte halves of the 32
// contraction bytes), each plane 16 core matrices of 8 rows x 16 bytes = the canonical K-major no-swizzle core-matrix order
// of the UMMA shared-memory descriptor.  Read with the contraction along the columns (K-major, 7 bulk copies of 4 KB per
// stage) it feeds T = K C; read with the contraction along the ROWS (MN-major: the same core matrices, instruction
// descriptor a_major = b_major = 1, one 5-D TMA box per operand and stage) it feeds K^T K -- the SYRK needs no transposed
// copy when the scale is matrix-wide.
// B-type operand (the symmetric C, row blocks of 64), stage-contiguous:
//   offset(r, k, p) = ((((r/64) * (Kd/32) + k/32) * 7 + p) * 2 + (k%32)/16) * 1024 + ((r%64)/8) * 128 + (r%8) * 16 + k%16
#pragma once
#include <cstdint>

#include "common.cuh"

namespace npgp {

constexpr int O8_NS = 7;                            // digits (bytes) per entry
constexpr int O8_BM = 128;                          // rows of an A-operand block (TMEM lanes)
constexpr int O8_BN = 64;                           // rows of a B-operand block (TMEM columns per accumulator)
constexpr int O8_KS = 32;                           // contraction bytes per pipeline stage = one MMA k-step
constexpr int O8_A_PLANE = O8_BM * O8_KS;           // 4096 bytes: one digit of one k-step of an A block
constexpr int O8_A_STAGE = O8_NS * O8_A_PLANE;      // 28672 bytes
constexpr int O8_B_STAGE = O8_NS * O8_BN * O8_KS;   // 14336 bytes
constexpr int O8_POISON = 1 << 20;                  // exponent marking a row / column that holds a non-finite entry
constexpr int O8_EMIN = -960;                       // exponents are clamped here so that 2^(54-e) stays a normal double
constexpr int O8_FRAC = 54;                         // y = rint(x 2^(O8_FRAC - e))
constexpr int O8_MAX_KD = 16384;                    // longest exact contraction per int32 accumulation (bound: 18724)
constexpr unsigned long long O8_BIAS = 0x0000808080808080ull;

__host__ __device__ inline long o8_digits_bytes(long rows, long Kd, int BR) {
  const long rpad = (rows + BR - 1) / BR * BR;
  return rpad * Kd * O8_NS;
}

// byte offset of the 16-byte vector holding columns k .. k+15 (k % 16 == 0) of row r, digit 0, in an A-type plane set
// with nks = Kd / 32 k-steps; digit p adds p * nks * 4096
__host__ __device__ inline long o8_a_offset(long r, int k, int nks) {
  return ((r / O8_BM) * O8_NS * nks + k / O8_KS) * (long)O8_A_PLANE + ((k % O8_KS) / 16) * (O8_BM * 16) +
         ((r % O8_BM) / 8) * 128 + (r % 8) * 16;
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

// exponent of a positive scale s (device scalar) bounding the entries: |x| <= s (1 + 1e-10) < 2^e
__device__ __forceinline__ int o8_exponent_of_scale(double s) {
  if (!(s > 0.0) || !(s < 1.7e308)) return O8_POISON;
  int e = 0;
  frexp(s * 1.0000000001, &e);
  return e < O8_EMIN ? O8_EMIN : e;
}

// Digit vectors of 16 consecutive contraction entries y[j] (|y| <= 2^54): digit p of the 16 entries is one 16-byte vector,
// stored at base + p * plane_stride.
__device__ __forceinline__ void o8_store_digits(const long long (&y)[16], int8_t* base, long plane_stride) {
  uint32_t lo[16], hi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const unsigned long long b = (unsigned long long)y[j] + O8_BIAS;
    lo[j] = (uint32_t)b;
    hi[j] = (uint32_t)(b >> 32);
  }
#pragma unroll
  for (int p = 0; p < O8_NS; ++p) {
    const int bi = (p < 3) ? (2 - p) : (6 - p);           // byte index inside hi (p < 3) or lo
    const uint32_t sel = (uint32_t)bi | ((uint32_t)(4 + bi) << 4);
    const uint32_t flip = (p == 0) ? 0u : 0x80808080u;    // byte - 128 for the six low digits
    uint32_t w[4];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      const uint32_t* src = (p < 3) ? hi : lo;
      const uint32_t t0 = __byte_perm(src[4 * g4], src[4 * g4 + 1], sel);
      const uint32_t t1 = __byte_perm(src[4 * g4 + 2], src[4 * g4 + 3], sel);
      w[g4] = __byte_perm(t0, t1, 0x5410) ^ flip;
    }
    *reinterpret_cast<uint4*>(base + (long)p * plane_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// The inverse: entry c (0..3) of the 4 consecutive columns held in word w[p] of each digit plane -> y as a double
__device__ __forceinline__ double o8_digits_to_double(const uint32_t (&w)[O8_NS], int c) {
  const uint32_t sel = (uint32_t)c | ((uint32_t)(4 + c) << 4);
  const uint32_t t0 = __byte_perm(w[6] ^ 0x80808080u, w[5] ^ 0x80808080u, sel);   // [b6, b5, ., .]
  const uint32_t t1 = __byte_perm(w[4] ^ 0x80808080u, w[3] ^ 0x80808080u, sel);   // [b4, b3, ., .]
  const uint32_t lo = __byte_perm(t0, t1, 0x5410);
  const uint32_t t2 = __byte_perm(w[2] ^ 0x80808080u, w[1] ^ 0x80808080u, sel);   // [b2, b1, ., .]
  // bytes 2, 3 = the sign-extended top digit (__byte_perm ignores the sign-replication bit of a selector nibble: measured)
  const int top = (int)(signed char)(w[0] >> (8 * c));
  const uint32_t hi = (t2 & 0xFFFFu) | ((uint32_t)top << 16);
  return __ll2double_rn((long long)((((unsigned long long)hi << 32) | lo) - O8_BIAS));
}

// y = rint(v * sc), sc = 2^(54-e), |v| 2^-e <= 1; a non-finite product becomes 0 (the row is poisoned through its exponent)
__device__ __forceinline__ long long o8_quantise(double v, double sc) {
  long long y = __double2ll_rn(v * sc);
  const long long lim = 1ll << 54;
  y = y > lim ? lim : y;
  y = y < -lim ? -lim : y;
  return y;
}

}  // namespace npgp
