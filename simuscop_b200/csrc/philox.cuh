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

}  // namespace ssc
