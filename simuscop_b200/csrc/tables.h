// tables.h -- host-side conversion of the reference's FP64 CDF tables into the exact
// integer-threshold tables the kernels search.
//
// The reference decides every draw with  r = ZF + (1-ZF)*(u/2^32);  first k: r <= cdf[k]
// (randIndx, lib/mydefine/MyDefine.cpp:176-184 with ThreadPool::randomDouble,
// lib/threadpool/ThreadPool.cpp:203-207).  r(u) is monotone non-decreasing in the 32-bit
// draw u, so {u : r(u) <= c} is a prefix [0, C) of the u range.  C is found here by binary
// search with the reference's own expression in FP64 (no contraction), which turns every
// FP64 compare on the device into a u32 compare with identical outcome for all 2^32 draws.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/simuscop.h"

namespace ssc {

// r(u) of randIndx (start = 2.2204e-16, end = 1), evaluated as the reference does.
double draw_real(uint32_t u, double start, double end);

// #{u in [0,2^32) : draw_real(u, ZF, 1) <= c}
uint64_t count_le(double c);

// Compressed CDF: strictly increasing inclusive thresholds T[i] with their symbols; the last
// entry always has T = 0xFFFFFFFF.  lookup(u) = sym[#{i : T[i] < u}].
struct CompressedCdf {
	std::vector<uint32_t> T;
	std::vector<uint16_t> sym;
};
CompressedCdf compress_cdf(const double* cdf, int ac);

// Substitution row (N = 4): {S0,S1,S2,base}; call = base + (u>S0) + (u>S1) + (u>S2).
struct SubRow { uint32_t s0, s1, s2, base; };
SubRow make_sub_row(const double* cdf4);

struct DeviceTablesHost {
	int N, K, B, Q, minQ, RL, paired, useCdf2, fixedInsert, minIS, nRows;
	// scalar event tests: p <= insertRate  <=>  u <= insT (if insEnable);  p2 < d  <=>  u <= delT (if delEnable)
	uint32_t insT, delT;
	int insEnable, delEnable;
	CompressedCdf isize, insLen, delLen;
	std::vector<SubRow> sub;        // [2 or 1][nRows*B]
	int qualPitch;                   // power of two >= longest compressed quality row
	std::vector<uint32_t> qualT;    // [N*N*B][qualPitch], padded with 0xFFFFFFFF
	std::vector<uint8_t> qualSym;   // [N*N*B][qualPitch], padded with the last symbol
	int maxQualRow;
	// rows with ref == call only (the ones nearly every base uses), padded to diagPitch (multiple of 4)
	int diagPitch;
	std::vector<uint32_t> qualDiagT;   // [N*B][diagPitch]
	std::vector<uint8_t> qualDiagSym;  // [N*B][diagPitch]
	uint8_t compLut;                 // 2 bits per code: complement code
	char baseChar[4];
	int8_t asciiCode[256];           // ASCII -> code 0..3, or 4 (non-ACGT)
};

// Returns an empty string on success, else an error message.
const char* build_tables(const ssc_profile_tables* t, DeviceTablesHost* out);

}  // namespace ssc
