/*
 * ssc_oracle_philox.h -- TEST INFRASTRUCTURE (oracle side). Not part of the product.
 *
 * Plain-C Philox4x32-10 (Salmon et al., SC'11; Random123 constants) and the
 * stream-address layout shared by the C oracle (ssc_oracle.c) and the
 * instrumented reference build (ref_shim/).  The product has its own,
 * independently written device implementation (simuscop_b200/csrc/philox.cuh);
 * both are pinned by the Random123 known-answer vectors in tests/.
 */
#ifndef SSC_ORACLE_PHILOX_H
#define SSC_ORACLE_PHILOX_H

#include <stdint.h>

enum { SSCO_STREAM_FRAG = 0, SSCO_STREAM_CYCLE = 1, SSCO_STREAM_LEN = 2, SSCO_STREAM_INSBASE = 3 };

static inline void ssco_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* address = (seed; pairID, mate, stream, blk, index) -> 4 words */
static inline void ssco_block(uint64_t seed, uint64_t pair, int mate, int stream, int blk, uint32_t index, uint32_t out[4]) {
	uint32_t ctr[4], key[2];
	ctr[0] = (uint32_t)pair;
	ctr[1] = (uint32_t)(pair >> 32);
	ctr[2] = ((uint32_t)mate << 28) | ((uint32_t)stream << 24) | (uint32_t)blk;
	ctr[3] = index;
	key[0] = (uint32_t)seed;
	key[1] = (uint32_t)(seed >> 32);
	ssco_philox4x32_10(ctr, key, out);
}

#endif
