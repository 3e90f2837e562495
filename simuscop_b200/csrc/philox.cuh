// philox.cuh -- Philox4x32-10 (Salmon et al. SC'11, Random123 constants) and the stream
// addressing of DESIGN.md.  Every draw site of the reference's per-read loop maps to one
// fixed word:   key = seed;  counter = (pairID_lo, pairID_hi, mate<<28 | stream<<24 | blk, index)
//   stream 0 (fragment), index = attempt : x0 start position (Segment.cpp:743), x1 insert size
//            (Profile.cpp:1491), x2 SE strand (Segment.cpp:766)
//   stream 1 (cycle),    index = j       : x0 insertion test, x1 deletion test at REFERENCE
//            position j (Profile.cpp:1560,1569); x2 substitution, x3 quality at OUTPUT
//            position j (Profile.cpp:1534/1547/1551, 1578/1583)
//   stream 2 (indel length), index = j   : x0 insertion length, x1 deletion length (:1520,:1524)
//   stream 3 (inserted bases), index = j, blk = b : bases 4b..4b+3 (:1564)
#pragma once
#include <cstdint>

namespace ssc {

enum { STREAM_FRAG = 0, STREAM_CYCLE = 1, STREAM_LEN = 2, STREAM_INSBASE = 3 };

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		c1 = (uint32_t)p1;
		c3 = (uint32_t)p0;
		c0 = n0;
		c2 = n2;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	u32x4 o;
	o.x = c0; o.y = c1; o.z = c2; o.w = c3;
	return o;
}

__host__ __device__ __forceinline__ u32x4 draw_block(uint64_t seed, uint64_t pair, int mate, int stream, int blk,
                                                      uint32_t index) {
	return philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32),
	                     ((uint32_t)mate << 28) | ((uint32_t)stream << 24) | (uint32_t)blk, index,
	                     (uint32_t)seed, (uint32_t)(seed >> 32));
}

#ifdef __CUDACC__
// ---- device-side forms shared by the generation kernel (gen_fast.cu) and the issue-rate microbenchmarks (floor.cu)
__device__ __forceinline__ void mulwide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
	asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

// Philox4x32-10 with the counter layout of philox.cuh::draw_block; rk = the 20 round keys
// (rk[2r] = seed_lo + r * 0x9E3779B9, rk[2r+1] = seed_hi + r * 0xBB67AE85), kernel parameters.
__device__ __forceinline__ u32x4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* rk) {
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint32_t lo0, hi0, lo1, hi1;
		mulwide(0xD2511F53u, c0, lo0, hi0);
		mulwide(0xCD9E8D57u, c2, lo1, hi1);
		c0 = hi1 ^ c1 ^ rk[2 * r];
		c2 = hi0 ^ c3 ^ rk[2 * r + 1];
		c1 = lo1;
		c3 = lo0;
	}
	u32x4 o;
	o.x = c0; o.y = c1; o.z = c2; o.w = c3;
	return o;
}

// NCH independent blocks (counter word 3 = base3 + 32*c), rounds interleaved across the blocks so
// that the dependent IMAD.WIDE -> LOP3 chains of the chunks overlap (ILP instead of occupancy).
template <int NCH>
__device__ __forceinline__ void philox_chunks(uint32_t pc0, uint32_t pc1, uint32_t pc2, uint32_t base3, const uint32_t* rk,
                                              uint32_t (&o0)[NCH], uint32_t (&o1)[NCH], uint32_t (&o2)[NCH], uint32_t (&o3)[NCH]) {
#pragma unroll
	for (int c = 0; c < NCH; c++) { o0[c] = pc0; o1[c] = pc1; o2[c] = pc2; o3[c] = base3 + 32u * c; }
#pragma unroll
	for (int r = 0; r < 10; r++) {
#pragma unroll
		for (int c = 0; c < NCH; c++) {
			uint32_t lo0, hi0, lo1, hi1;
			mulwide(0xD2511F53u, o0[c], lo0, hi0);
			mulwide(0xCD9E8D57u, o2[c], lo1, hi1);
			o0[c] = hi1 ^ o1[c] ^ rk[2 * r];
			o2[c] = hi0 ^ o3[c] ^ rk[2 * r + 1];
			o1[c] = lo1;
			o3[c] = lo0;
		}
	}
}

#endif  // __CUDACC__

}  // namespace ssc
