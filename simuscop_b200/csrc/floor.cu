// floor.cu -- issue-rate microbenchmarks: the measured ceiling under which generate_slots_kernel (gen_fast.cu) works.
//
// The generation path is bound by instruction issue, not by HBM: the reference draws four uniform numbers per simulated
// base (Profile.cpp:1560, 1569, 1534/1547, 1578), i.e. one Philox4x32-10 block per base.  These two kernels measure, on the
// GPU the bench runs on, what that costs when nothing else is done:
//   mode 0  Philox only: the generation kernel's launch shape (one 1024-thread CTA per SM, a warp takes tickets of 32 pairs
//           from an atomic counter, lane = cycle, NCH interleaved blocks per mate, round keys as kernel parameters), every
//           draw xor-folded into one word per warp, one store per warp at the end;
//   mode 1  the same plus the fast per-base path of an indel-free read on synthetic tables of the real shapes (context cut
//           out of a packed window, context LUT, substitution row LDS.128 + three compare-adds, three-step quality search,
//           symbol + base character in one 16-bit load) and the two byte stores per base into per-ticket blobs: no
//           prologue, no header, no candidate tests, no indel path, no pass 2.
// bench.py reports them as roofline.issue; neither is part of the product path.
#include <cuda_runtime.h>
#include <cstdint>

#include "kernels.h"
#include "philox.cuh"

namespace ssc {

namespace {

constexpr int FL_QROW = 68;        // bytes of a shared-memory quality row (gen_fast.cu F_QROW)
constexpr int FL_ROWS = 84;        // k-mer rows (K = 3)
constexpr int FL_BINS = 50;
constexpr int FL_PITCH = 51;       // odd entry pitch of a substitution row
constexpr int FL_QBINS = 56;       // bins per (ref, call) block of the quality image

struct FloorParams {
	uint32_t rk[20];
	uint32_t one;
	int RL, nTiles, nPairs;
	unsigned int* ticket;
	uint32_t* sink;            // one word per warp
	uint8_t* blobs;            // mode 1: blob scratch, blobPitch bytes per ticket, file 2 behind file 1
	uint32_t blobPitch, file2Off;
};

__device__ __forceinline__ uint32_t fl_lds_u32(uint32_t addr) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t fl_lds_u16(uint32_t addr) { uint32_t v; asm("{\n\t.reg .u16 h;\n\tld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void fl_fadd_gt(uint32_t& acc, uint32_t a, uint32_t b, uint32_t inc, uint32_t one) {
	asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %4, %3, %0;\n\t}" : "+r"(acc) : "r"(a), "r"(b), "r"(inc), "r"(one));
}
__device__ __forceinline__ void fl_fadd_lt(uint32_t& acc, uint32_t a, uint32_t b, uint32_t inc, uint32_t one) {
	asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %4, %3, %0;\n\t}" : "+r"(acc) : "r"(a), "r"(b), "r"(inc), "r"(one));
}

template <int NCH, int MODE>
__global__ void __launch_bounds__(FG_THREADS, 1) issue_floor_kernel(const __grid_constant__ FloorParams P) {
	extern __shared__ __align__(16) uint8_t smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int RL = P.RL;
	// shared image (mode 1): substitution rows of both mates | quality rows | context LUT | one packed window per warp
	uint4* s_sub = (uint4*)smem;
	uint8_t* s_qual = smem + 2 * FL_ROWS * FL_PITCH * 16;
	uint16_t* s_lut = (uint16_t*)(s_qual + 16 * FL_QBINS * FL_QROW);
	uint32_t* s_win = (uint32_t*)(s_lut + 128) + warp * 32;
	const uint32_t qualBaseS = (uint32_t)__cvta_generic_to_shared(s_qual);
	const uint32_t qstride = FL_QBINS * FL_QROW;
	if (MODE == 1) {
		uint32_t lcg = 2463534242u + threadIdx.x;
		for (int i = threadIdx.x; i < 2 * FL_ROWS * FL_PITCH; i += FG_THREADS) {
			const int row = (i / FL_PITCH) % FL_ROWS;
			lcg = lcg * 1664525u + 1013904223u;
			// a typical row: the three thresholds sit in the top percent of the draw range (substitutions are rare)
			const uint32_t s0 = 0xFD000000u + (lcg >> 10);
			s_sub[i] = make_uint4(s0, s0 + 0x00400000u, s0 + 0x00800000u, qualBaseS + (uint32_t)((row & 3) * 4) * qstride);
		}
		for (int i = threadIdx.x; i < 16 * FL_QBINS * 8; i += FG_THREADS) {
			const int r = i >> 3, k = i & 7;
			uint32_t* dst = (uint32_t*)(s_qual + r * FL_QROW + k * 8);
			dst[0] = k < 7 ? (uint32_t)(k + 1) * 0x20000000u + (uint32_t)(r * 2654435761u >> 8) : 0xFFFFFFFFu;
			dst[1] = (uint32_t)('#' + k * 5) | ((uint32_t)"ACTG"[(r / FL_QBINS) & 3] << 8);
		}
		for (int i = threadIdx.x; i < 128; i += FG_THREADS) {
			const uint32_t b2 = (i >> 4) & 3, b1 = (i >> 2) & 3, b0 = i & 3;
			s_lut[i] = (uint16_t)((20u + 16u * b0 + 4u * b1 + b2) * FL_PITCH);
		}
		for (int i = lane; i < 32; i += 32) s_win[i] = (uint32_t)(warp * 97 + i) * 2654435761u;
		__syncthreads();
	}
	uint32_t binOf[NCH];
#pragma unroll
	for (int c = 0; c < NCH; c++) {
		int j = c * 32 + lane;
		if (j > RL - 1) j = RL - 1;
		binOf[c] = (uint32_t)((j * FL_BINS) / RL);
	}
	const int jLast = (NCH - 1) * 32 + lane;
	const uint32_t one = P.one;
	uint32_t fold = 0;
	while (true) {
		int chunk = 0;
		if (lane == 0) chunk = (int)atomicAdd(P.ticket, 1u);
		chunk = __shfl_sync(0xffffffffu, chunk, 0);
		if (chunk >= P.nTiles) break;
		const uint32_t slot0 = (uint32_t)chunk * FG_CHUNK;
		const int count = (int)(slot0 + FG_CHUNK < (uint32_t)P.nPairs ? FG_CHUNK : (uint32_t)P.nPairs - slot0);
		const uint32_t blobBase = (uint32_t)chunk * P.blobPitch;
		uint32_t posA = blobBase, posB = blobBase + P.file2Off;
#pragma unroll 1
		for (int p = 0; p < count; p++) {
			const uint32_t c0 = slot0 + (uint32_t)p, c1 = 0u;
#pragma unroll 1
			for (int mate = 0; mate < 2; mate++) {
				const uint32_t c2cyc = ((uint32_t)mate << 28) | ((uint32_t)STREAM_CYCLE << 24);
				uint32_t x0[NCH], x1[NCH], x2[NCH], x3[NCH];
				philox_chunks<NCH>(c0, c1, c2cyc, (uint32_t)lane, P.rk, x0, x1, x2, x3);
				if (MODE == 0) {
#pragma unroll
					for (int c = 0; c < NCH; c++) { fold ^= x0[c] ^ x1[c]; fold ^= x2[c] ^ x3[c]; }
				} else {
#pragma unroll
					for (int c = 0; c < NCH; c++) fold ^= x0[c] ^ x1[c];
					const bool rev = mate == 1;
					const int dOff = (int)(c0 & 15u) + 32, H = 30;
					const int rel0 = rev ? (dOff + RL - 1 - lane) : (dOff + lane - 2);
					const uint32_t dsh = (uint32_t)(rel0 & 15) * 2u;
					const uint32_t* dptr = s_win + (rel0 >> 4);
					const int dstep = rev ? -2 : 2;
					const uint8_t* lutB = (const uint8_t*)s_lut + (rev ? 128 : 0);
					const uint8_t* subMB = (const uint8_t*)(s_sub + (rev ? FL_ROWS * FL_PITCH : 0));
					uint8_t* st1 = P.blobs + posA + H + lane;
					uint8_t* st2 = st1 + RL + 3;
#pragma unroll
					for (int c = 0; c < NCH; c++) {
						const uint32_t v6 = __funnelshift_r(dptr[c * dstep], dptr[c * dstep + 1], dsh) & 63u;
						const uint32_t rowIdx = *(const uint16_t*)(lutB + v6 * 2u);
						const uint4 sr = *(const uint4*)(subMB + (rowIdx + binOf[c]) * 16u);
						uint32_t acc = sr.w;
						fl_fadd_gt(acc, x2[c], sr.x, qstride, one);
						fl_fadd_gt(acc, x2[c], sr.y, qstride, one);
						fl_fadd_gt(acc, x2[c], sr.z, qstride, one);
						uint32_t qa = binOf[c] * (uint32_t)FL_QROW + acc;
						fl_fadd_lt(qa, fl_lds_u32(qa + 24), x3[c], 32u, one);
						fl_fadd_lt(qa, fl_lds_u32(qa + 8), x3[c], 16u, one);
						fl_fadd_lt(qa, fl_lds_u32(qa), x3[c], 8u, one);
						const uint32_t q = fl_lds_u16(qa + 4);
						if (c < NCH - 1 || jLast < RL) {
							st1[c * 32] = (uint8_t)(q >> 8);
							st2[c * 32] = (uint8_t)q;
						}
					}
					posA += (uint32_t)(H + 2 * RL + 4);
					const uint32_t t = posA; posA = posB; posB = t;
				}
			}
		}
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) fold ^= __shfl_xor_sync(0xffffffffu, fold, d);
	if (lane == 0) P.sink[blockIdx.x * FG_WORKERS + warp] = fold;
}

template <int NCH>
cudaError_t launch_floor_nch(int mode, const FloorParams& P, int grid, cudaStream_t s) {
	if (mode == 0) {
		issue_floor_kernel<NCH, 0><<<grid, FG_THREADS, 0, s>>>(P);
		return cudaGetLastError();
	}
	const int smem = 2 * FL_ROWS * FL_PITCH * 16 + 16 * FL_QBINS * FL_QROW + 256 + FG_WORKERS * 128;
	auto kern = issue_floor_kernel<NCH, 1>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (e != cudaSuccess) return e;
	kern<<<grid, FG_THREADS, smem, s>>>(P);
	return cudaGetLastError();
}

}  // namespace

// One launch over nPairs pairs of read length RL.  scratch: >= 4 + 4 * grid * FG_WORKERS bytes (ticket counter + one word per
// warp); blobs (mode 1): 2 * ceil(nPairs / FG_CHUNK) * blobPitch bytes.
cudaError_t launch_issue_floor(int mode, int RL, int64_t nPairs, uint64_t seed, int grid, uint32_t* scratch, uint8_t* blobs,
                               uint32_t blobPitch, uint32_t file2Off, cudaStream_t stream) {
	FloorParams P;
	for (int r = 0; r < 10; r++) {
		P.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
		P.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
	}
	P.one = 1; P.RL = RL; P.nPairs = (int)nPairs; P.nTiles = (int)((nPairs + FG_CHUNK - 1) / FG_CHUNK);
	P.ticket = scratch; P.sink = scratch + 1; P.blobs = blobs; P.blobPitch = blobPitch; P.file2Off = file2Off;
	cudaError_t e = cudaMemsetAsync(scratch, 0, 4, stream);
	if (e != cudaSuccess) return e;
	const int nch = (RL + 31) / 32;
	if (nch <= 2) return launch_floor_nch<2>(mode, P, grid, stream);
	if (nch == 3) return launch_floor_nch<3>(mode, P, grid, stream);
	if (nch == 4) return launch_floor_nch<4>(mode, P, grid, stream);
	return launch_floor_nch<5>(mode, P, grid, stream);
}

}  // namespace ssc
