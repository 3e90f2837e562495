// gen_fast.cu -- the production generation kernel (N = 4, K = 3, RL <= 160, tables in shared memory).
//
// Same algorithm and bytes as generate_kernel (kernels.cu), engineered for instruction issue,
// which -- not HBM -- bounds this path (4 uniform draws per base = one Philox4x32-10 block per
// lane per cycle):
//   * one warp per pair, lane = sequencing cycle, both mates; FG_WORKERS independent warps per
//     CTA, one CTA per SM (148), static assignment of consecutive pairs to consecutive warps;
//   * Philox rounds as mul.wide.u32 (IMAD.WIDE) + LOP3;
//   * the k-mer context of cycle j is cut straight out of the 2-bit packed haplotype window
//     (two LDS + one funnel shift for three bases, a 128-entry LUT turns it into the table row),
//     so the common no-indel read never materialises its bases;
//   * indel candidates are only OR-ed per lane; one vote per read decides fast vs. slow path,
//     the slow path (about 15 % of reads) is a compact non-unrolled routine;
//   * warps never wait for each other: every record is written to its own fixed-pitch slot in an
//     HBM scratch slab (pass 1); a bandwidth-bound second kernel scans the record lengths
//     (decoupled look-back over 256-pair tiles) and copies the records to their exact byte offset
//     with 16-byte stores (pass 2), so the final slab is dense, ordered and byte-identical to the
//     reference's file.  HBM traffic is ~3x the algorithmic bytes, at < 5 % of HBM bandwidth,
//     in exchange for removing every barrier from the issue-bound generation pass.
#include <cuda_runtime.h>
#include <cstdint>

#include "device_types.h"
#include "kernels.h"
#include "philox.cuh"

#ifdef PHILOX_UNROLL
static constexpr int SSC_PHILOX_UNROLL = PHILOX_UNROLL;
#else
static constexpr int SSC_PHILOX_UNROLL = 2;   // rounds per loop trip: keeps the hot loop inside the instruction cache
#endif

namespace ssc {

#define ST_A (1ull << 62)
#define ST_P (2ull << 62)
#define ST_MASK (3ull << 62)

static constexpr int F_SRC_CAP = 256;           // longest read after indels
static constexpr int F_EV_MAX = 32;
static constexpr int F_INS_CAP = 128;
static constexpr int F_WIN_WORDS = 32;          // 16 data words + 9 mask words (+pad)

struct FastLayout {
	int sub, qualT, qualSym, isizeT, isizeSym, insT, insSym, delT, delSym, lut, warp, total;
	int w_stage, w_src, w_ev, w_insb, w_win, perWarp;
};

__host__ __device__ inline FastLayout fast_layout(int nSubTotal, int nQual, int nIsize, int nIns, int nDel) {
	FastLayout L;
	int o = 0;
	L.sub = o; o += nSubTotal * 16;
	L.qualT = o; o += nQual * 4;
	L.qualSym = o; o += (nQual + 15) / 16 * 16;
	L.isizeT = o; o += (nIsize * 4 + 15) / 16 * 16;
	L.isizeSym = o; o += (nIsize * 2 + 15) / 16 * 16;
	L.insT = o; o += (nIns * 4 + 15) / 16 * 16;
	L.insSym = o; o += (nIns * 2 + 15) / 16 * 16;
	L.delT = o; o += (nDel * 4 + 15) / 16 * 16;
	L.delSym = o; o += (nDel * 2 + 15) / 16 * 16;
	L.lut = o; o += 128 * 2;
	int w = 0;
	L.w_stage = w; w += 2 * 160 * 4;            // x2/x3 of the current mate, saved for the slow path
	L.w_src = w; w += F_SRC_CAP;
	L.w_ev = w; w += F_EV_MAX * 4;
	L.w_insb = w; w += F_INS_CAP;
	L.w_win = w; w += F_WIN_WORDS * 4;
	L.perWarp = w;
	L.warp = o; o += FG_WORKERS * L.perWarp;
	L.total = o;
	return L;
}

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void mulwide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
	asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

// acc += (a > b) for unsigned a, b: one ISETP + one predicated IADD
__device__ __forceinline__ void add_gt(uint32_t& acc, uint32_t a, uint32_t b) {
	asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(acc) : "r"(a), "r"(b));
}
// acc += inc if a < b (unsigned)
__device__ __forceinline__ void add_lt(uint32_t& acc, uint32_t a, uint32_t b, uint32_t inc) {
	asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(acc) : "r"(a), "r"(b), "r"(inc));
}

// Philox4x32-10 with the counter layout of philox.cuh::draw_block
__device__ __forceinline__ u32x4 philox_fast(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint32_t lo0, hi0, lo1, hi1;
		mulwide(0xD2511F53u, c0, lo0, hi0);
		mulwide(0xCD9E8D57u, c2, lo1, hi1);
		c0 = hi1 ^ c1 ^ k0;
		c2 = hi0 ^ c3 ^ k1;
		c1 = lo1;
		c3 = lo0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	u32x4 o;
	o.x = c0; o.y = c1; o.z = c2; o.w = c3;
	return o;
}

// NCH independent blocks (counter word 3 = base3 + 32*c), rounds interleaved across the blocks so
// that the dependent IMAD.WIDE -> LOP3 chains of the chunks overlap (ILP instead of occupancy).
template <int NCH>
__device__ __forceinline__ void philox_chunks(uint32_t pc0, uint32_t pc1, uint32_t pc2, uint32_t base3, uint32_t k0, uint32_t k1,
                                              uint32_t (&o0)[NCH], uint32_t (&o1)[NCH], uint32_t (&o2)[NCH], uint32_t (&o3)[NCH]) {
#pragma unroll
	for (int c = 0; c < NCH; c++) { o0[c] = pc0; o1[c] = pc1; o2[c] = pc2; o3[c] = base3 + 32u * c; }
#pragma unroll (SSC_PHILOX_UNROLL)
	for (int r = 0; r < 10; r++) {
#pragma unroll
		for (int c = 0; c < NCH; c++) {
			uint32_t lo0, hi0, lo1, hi1;
			mulwide(0xD2511F53u, o0[c], lo0, hi0);
			mulwide(0xCD9E8D57u, o2[c], lo1, hi1);
			o0[c] = hi1 ^ o1[c] ^ k0;
			o2[c] = hi0 ^ o3[c] ^ k1;
			o1[c] = lo1;
			o3[c] = lo0;
		}
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
}

__device__ __forceinline__ long long f_draw_pos(uint32_t u, int spos, int epos) {
	double frac = __dmul_rn((double)u, 2.3283064365386962890625e-10);
	double v = __dadd_rn((double)spos, __dmul_rn((double)((long long)epos + 1 - spos), frac));
	return (long long)v;
}

// warp-cooperative lookup in a compressed CDF held in shared memory (n <= 1024): sym[#{i : T[i] < u}]
__device__ __forceinline__ int coop_lookup(const uint32_t* T, const uint16_t* sym, int n, uint32_t u, int lane) {
	const int stride = (n + 31) >> 5;
	int i1 = (lane + 1) * stride - 1;
	if (i1 > n - 1) i1 = n - 1;
	const int blk = __popc(__ballot_sync(0xffffffffu, T[i1] < u && (lane + 1) * stride <= n));
	const int i2 = blk * stride + lane;
	const bool v2 = lane < stride && i2 < n && T[i2] < u;
	const int cnt = blk * stride + __popc(__ballot_sync(0xffffffffu, v2));
	return (int)sym[cnt < n ? cnt : n - 1];
}

// uniform binary search (all lanes the same u), small tables
__device__ __forceinline__ int uni_lookup(const uint32_t* T, const uint16_t* sym, int n, uint32_t u) {
	int lo = 0, len = n - 1;
	while (len > 0) {
		int half = len >> 1;
		if (T[lo + half] < u) { lo += half + 1; len -= half + 1; } else len = half;
	}
	return (int)sym[lo];
}

__device__ __forceinline__ int f_ndigits(uint32_t v) {
	int n = 1;
	n += v >= 10u; n += v >= 100u; n += v >= 1000u; n += v >= 10000u; n += v >= 100000u;
	n += v >= 1000000u; n += v >= 10000000u; n += v >= 100000000u; n += v >= 1000000000u;
	return n;
}

__constant__ uint32_t c_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};

struct QualTabs {
	const uint32_t* T; const uint8_t* sym;            // QP == 8: shared, pitch 8
	const uint32_t* diagT; const uint8_t* diagSym;    // QP == 2: shared, ref == call rows
	const uint32_t* gT; const uint8_t* gSym;          // full table in global memory
	int pitch, diagPitch;
};

struct WarpCtx {
	int B, qualPitch, minQ, RL, nInsLen, nDelLen, nBasesM1, mDelta;
	uint32_t baseChars, compLut;
	const uint4* sub;          // shared: table of the current mate
	QualTabs q;
	const uint32_t* insT; const uint16_t* insSym;
	const uint32_t* delT; const uint16_t* delSym;
	const uint32_t* win;       // shared window: data words [0..16), mask words [16..25)
	uint8_t* src; uint32_t* ev; uint8_t* insb;
	uint32_t k0, k1, c0, c1;   // Philox key and pair counter words
	int lane;
};

// substitution + quality for one output base; returns (char | qual << 8)
template <int QP> __device__ __forceinline__ uint32_t qual_lookup(const QualTabs& q, uint32_t cur, uint32_t call, uint32_t binIdx, int B, uint32_t u3);

template <int QP>
__device__ __forceinline__ uint32_t call_base(const WarpCtx& w, uint32_t cur, int row, bool bad, bool curN, int binIdx,
                                              uint32_t u2, uint32_t u3) {
	int call;
	if (bad) call = curN ? -1 : (int)cur;
	else {
		const uint4 s = w.sub[row * w.B + binIdx];
		call = (int)s.w + (u2 > s.x) + (u2 > s.y) + (u2 > s.z);
	}
	uint32_t ch, q;
	if (call < 0) { ch = 'N'; q = (uint32_t)w.minQ + __umulhi(20u, u3); }    // randomInteger(33, 53), Profile.cpp:1583
	else {
		ch = __byte_perm(w.baseChars, 0, 0x4440 | call);
		q = qual_lookup<QP>(w.q, cur, (uint32_t)call, (uint32_t)binIdx, w.B, u3);
	}
	return ch | (q << 8);
}

// template base code (0..3, 4 = non-ACGT) of read position j, from the shared window
__device__ __forceinline__ uint32_t window_code(const WarpCtx& w, int relBase /* base index relative to window word 0 */,
                                                bool rev) {
	const uint32_t d = w.win[relBase >> 4];
	uint32_t code = (d >> ((relBase & 15) * 2)) & 3u;
	if (rev) code = (w.compLut >> (2 * code)) & 3u;
	const int relM = relBase + w.mDelta;
	const uint32_t mk = w.win[16 + (relM >> 5)];
	if ((mk >> (relM & 31)) & 1u) code = 4u;
	return code;
}

// Slow path of Profile::predict (a read with at least one indel candidate): compact, not unrolled.
// evbits: per lane, bit 2c = insertion test hit at cycle 32c+lane, bit 2c+1 = deletion test hit.
// Writes bases/quals into stage[H ..]; returns m.
template <int QP>
__device__ __forceinline__ int slow_read(const WarpCtx& w, uint32_t evbits, int mate, bool rev, int relFirst,
                                      uint8_t* stage, int H, unsigned int* errorFlags, const uint32_t* xsave, int nSaved) {
	const int RL = w.RL, lane = w.lane;
	const int chunksRL = (RL + 31) >> 5;
	const uint32_t c2cyc = ((uint32_t)mate << 28) | ((uint32_t)STREAM_CYCLE << 24);
	const uint32_t c2len = ((uint32_t)mate << 28) | ((uint32_t)STREAM_LEN << 24);
	const uint32_t c2ins = ((uint32_t)mate << 28) | ((uint32_t)STREAM_INSBASE << 24);
	int nEv = 0, insTotal = 0, indelLength = 0, skipUntil = 0;
	bool tooMany = false;
	for (int c = 0; c < chunksRL; c++) {
		const uint32_t insMask = __ballot_sync(0xffffffffu, (evbits >> (2 * c)) & 1u);
		const uint32_t delMask = __ballot_sync(0xffffffffu, (evbits >> (2 * c + 1)) & 1u);
		uint32_t mask = insMask | delMask;
		while (mask) {
			const int bit = __ffs(mask) - 1;
			mask &= mask - 1;
			const int j = c * 32 + bit;
			if (j < skipUntil) continue;
			const u32x4 lb = philox_fast(w.c0, w.c1, c2len, (uint32_t)j, w.k0, w.k1);
			if ((insMask >> bit) & 1u) {
				const int Lk = uni_lookup(w.insT, w.insSym, w.nInsLen, lb.x);          // Profile::getInsertLen
				if (Lk > 0) {
					if (nEv >= F_EV_MAX || insTotal + Lk > F_INS_CAP) { tooMany = true; break; }
					for (int i = lane; i < Lk; i += 32) {                                  // Profile.cpp:1563-1566
						const u32x4 bb = philox_fast(w.c0, w.c1, c2ins | (uint32_t)(i >> 2), (uint32_t)j, w.k0, w.k1);
						const uint32_t ws = (i & 3) == 0 ? bb.x : (i & 3) == 1 ? bb.y : (i & 3) == 2 ? bb.z : bb.w;
						w.insb[insTotal + i] = (uint8_t)__umulhi((uint32_t)w.nBasesM1, ws);
					}
					if (lane == 0) w.ev[nEv] = (uint32_t)j | ((uint32_t)Lk << 12) | ((uint32_t)insTotal << 20) | (1u << 31);
					nEv++; insTotal += Lk; indelLength += Lk;
				}
			} else {
				int Lk = uni_lookup(w.delT, w.delSym, w.nDelLen, lb.y);                  // Profile::getDelLen
				if (Lk > RL - j) Lk = RL - j;                                             // Profile.cpp:1613
				if (Lk > 0) {
					if (nEv >= F_EV_MAX) { tooMany = true; break; }
					if (lane == 0) w.ev[nEv] = (uint32_t)j | ((uint32_t)Lk << 12);
					nEv++; indelLength -= Lk; skipUntil = j + Lk;
				}
			}
		}
	}
	if (RL + indelLength < 50) { nEv = 0; indelLength = 0; }                             // Profile.cpp:1627-1634
	int m = RL + indelLength;
	if (tooMany || m > F_SRC_CAP) {
		if (lane == 0) atomicOr(errorFlags, tooMany ? 4u : 2u);
		nEv = 0; m = RL;
	}
	__syncwarp();
	// source sequence: template bases moved to their output positions, inserted bases after their base
	for (int c = 0; c < chunksRL; c++) {
		const int j = c * 32 + lane;
		if (j < RL) {
			const uint32_t code = window_code(w, rev ? relFirst - j : relFirst + j, rev);
			int shift = 0; bool dropped = false;
			for (int k = 0; k < nEv; k++) {
				const uint32_t ev = w.ev[k];
				const int ej = (int)(ev & 0xfffu), el = (int)((ev >> 12) & 0xffu);
				if (ev >> 31) { if (ej < j) shift += el; }
				else { if (j >= ej && j < ej + el) dropped = true; else if (j >= ej + el) shift -= el; }
			}
			if (!dropped) w.src[j + shift] = (uint8_t)code;
		}
	}
	int cum = 0;
	for (int k = 0; k < nEv; k++) {
		const uint32_t ev = w.ev[k];
		const int ej = (int)(ev & 0xfffu), el = (int)((ev >> 12) & 0xffu);
		if (ev >> 31) {
			const int io = (int)((ev >> 20) & 0x7ffu);
			for (int i = lane; i < el; i += 32) w.src[ej + cum + 1 + i] = w.insb[io + i];
			cum += el;
		} else cum -= el;
	}
	__syncwarp();
	const uint32_t inv = (m > 1) ? (0xffffffffu / (uint32_t)m + 1u) : 0xffffffffu;
	const int chunksM = (m + 31) >> 5;
	for (int c = 0; c < chunksM; c++) {
		const int j = c * 32 + lane;
		if (j < m) {
			uint32_t u2, u3;
			if (j < nSaved) { u2 = xsave[j]; u3 = xsave[160 + j]; }
			else { const u32x4 blk = philox_fast(w.c0, w.c1, c2cyc, (uint32_t)j, w.k0, w.k1); u2 = blk.z; u3 = blk.w; }
			const uint32_t cur = w.src[j];
			const uint32_t p1 = j >= 1 ? w.src[j - 1] : 0u;
			const uint32_t p2 = j >= 2 ? w.src[j - 2] : 0u;
			const bool bad = ((cur | p1 | p2) & 4u) != 0;
			const int row = j >= 2 ? (int)(20u + 16u * (p2 & 3u) + 4u * (p1 & 3u) + (cur & 3u))
			                       : j == 1 ? (int)(4u + 4u * (p1 & 3u) + (cur & 3u)) : (int)(cur & 3u);
			const int binIdx = (int)__umulhi((uint32_t)(j * w.B), inv);
			const uint32_t r = call_base<QP>(w, cur & 3u, row, bad, (cur & 4u) != 0, binIdx, u2, u3);
			stage[H + j] = (uint8_t)r;
			stage[H + m + 3 + j] = (uint8_t)(r >> 8);
		}
	}
	return m;
}

// Quality lookup.  Rows hold ascending inclusive thresholds padded with 0xFFFFFFFF; symbol index = #{i : T[i] < u}.
//   QP == 8: all 16*B rows in shared memory, pitch 8 (XTen-like profiles), three unrolled steps;
//   QP == 2: only the ref == call rows in shared memory (pitch = live symbols rounded up to 4, generic branch-free
//            lower bound); the rare substituted bases go to the full table in global memory (L2);
//   QP == 0: full table in global memory.
template <int QP>
__device__ __forceinline__ uint32_t qual_lookup(const QualTabs& q, uint32_t cur, uint32_t call, uint32_t binIdx, int B, uint32_t u3) {
	if (QP == 8) {
		const uint32_t qrow = (cur * 4u + call) * (uint32_t)B + binIdx;
		const uint32_t* qt = q.T + qrow * 8;
		uint32_t k = 0;
		add_lt(k, qt[3], u3, 4u);
		add_lt(k, qt[k + 1], u3, 2u);
		add_lt(k, qt[k], u3, 1u);
		return q.sym[qrow * 8 + k];
	}
	if (QP == 2 && cur == call) {
		const uint32_t base = (cur * (uint32_t)B + binIdx) * (uint32_t)q.diagPitch;
		const uint32_t* qt = q.diagT + base;
		uint32_t k = 0;
		for (int len = q.diagPitch; len > 1;) {
			const int half = len >> 1;
			add_lt(k, qt[k + half - 1], u3, (uint32_t)half);
			len -= half;
		}
		return q.diagSym[base + k];
	}
	const uint32_t qrow = (cur * 4u + call) * (uint32_t)B + binIdx;
	const uint32_t* qt = q.gT + qrow * q.pitch;
	int k = 0;
	for (int s = q.pitch >> 1; s > 0; s >>= 1) if (qt[k + s - 1] < u3) k += s;
	return q.gSym[qrow * q.pitch + k];
}

// ---------------------------------------------------------------------------------------------
// pass 1: generation into fixed-pitch slots
// ---------------------------------------------------------------------------------------------
template <int NCH, int QP>
__global__ void __launch_bounds__(FG_THREADS, 1) generate_slots_kernel(const GenParams P) {
	extern __shared__ __align__(16) uint8_t smem[];
	const DevTables& t = P.t;
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int nSubTotal = t.nSub * (t.useCdf2 ? 2 : 1);
	const int nQual = QP == 8 ? t.nQualRows * 8 : (QP == 2 ? 4 * t.B * t.qualDiagPitch : 0);
	const FastLayout L = fast_layout(nSubTotal, nQual, t.nIsize, t.nInsLen, t.nDelLen);

	uint4* s_sub = (uint4*)(smem + L.sub);
	uint32_t* s_qualT = (uint32_t*)(smem + L.qualT);
	uint8_t* s_qualSym = smem + L.qualSym;
	uint32_t* s_isizeT = (uint32_t*)(smem + L.isizeT);
	uint16_t* s_isizeSym = (uint16_t*)(smem + L.isizeSym);
	uint32_t* s_insT = (uint32_t*)(smem + L.insT);
	uint16_t* s_insSym = (uint16_t*)(smem + L.insSym);
	uint32_t* s_delT = (uint32_t*)(smem + L.delT);
	uint16_t* s_delSym = (uint16_t*)(smem + L.delSym);
	uint16_t* s_lut = (uint16_t*)(smem + L.lut);

#pragma unroll 1
	for (int i = threadIdx.x; i < nSubTotal; i += FG_THREADS) s_sub[i] = t.sub[i];
	{
		const uint32_t* srcT = QP == 8 ? t.qualT : t.qualDiagT;
		const uint8_t* srcS = QP == 8 ? t.qualSym : t.qualDiagSym;
#pragma unroll 1
		for (int i = threadIdx.x; i < nQual / 4; i += FG_THREADS) ((uint4*)s_qualT)[i] = ((const uint4*)srcT)[i];
#pragma unroll 1
		for (int i = threadIdx.x; i < nQual / 16; i += FG_THREADS) ((uint4*)s_qualSym)[i] = ((const uint4*)srcS)[i];
	}
#pragma unroll 1
	for (int i = threadIdx.x; i < t.nIsize; i += FG_THREADS) { s_isizeT[i] = t.isizeT[i]; s_isizeSym[i] = t.isizeSym[i]; }
#pragma unroll 1
	for (int i = threadIdx.x; i < t.nInsLen; i += FG_THREADS) { s_insT[i] = t.insLenT[i]; s_insSym[i] = t.insLenSym[i]; }
#pragma unroll 1
	for (int i = threadIdx.x; i < t.nDelLen; i += FG_THREADS) { s_delT[i] = t.delLenT[i]; s_delSym[i] = t.delLenSym[i]; }
	if (threadIdx.x < 128) {
		// context LUT: index = dir << 6 | b0 | b1 << 2 | b2 << 4 (three consecutive store bases)
		// forward: (b0,b1,b2) = (j-2, j-1, j); reverse: (b0,b1,b2) = raw bases of (j, j-1, j-2)
		const int v = threadIdx.x & 63, dir = threadIdx.x >> 6;
		const uint32_t b0 = v & 3, b1 = (v >> 2) & 3, b2 = (v >> 4) & 3;
		uint32_t cur, p1, p2;
		if (!dir) { p2 = b0; p1 = b1; cur = b2; }
		else { cur = (t.compLut >> (2 * b0)) & 3u; p1 = (t.compLut >> (2 * b1)) & 3u; p2 = (t.compLut >> (2 * b2)) & 3u; }
		s_lut[threadIdx.x] = (uint16_t)((20u + 16u * p2 + 4u * p1 + cur) | (cur << 8) | (p1 << 10));
	}
	__syncthreads();

	const int RL = t.RL, B = t.B;
	const int nMates = t.paired ? 2 : 1;

	uint8_t* wbase = smem + L.warp + warp * L.perWarp;
	WarpCtx w;
	w.B = t.B; w.qualPitch = t.qualPitch; w.minQ = t.minQ; w.RL = t.RL; w.nInsLen = t.nInsLen; w.nDelLen = t.nDelLen;
	w.nBasesM1 = t.N - 1; w.mDelta = 0; w.baseChars = t.baseChars; w.compLut = t.compLut;
	w.q.T = s_qualT; w.q.sym = s_qualSym; w.q.diagT = s_qualT; w.q.diagSym = s_qualSym;
	w.q.gT = t.qualT; w.q.gSym = t.qualSym; w.q.pitch = t.qualPitch; w.q.diagPitch = t.qualDiagPitch;
	w.insT = s_insT; w.insSym = s_insSym; w.delT = s_delT; w.delSym = s_delSym;
	w.win = (const uint32_t*)(wbase + L.w_win);
	uint32_t* s_win = (uint32_t*)(wbase + L.w_win);
	uint32_t* s_xsave = (uint32_t*)(wbase + L.w_stage);
	w.src = wbase + L.w_src; w.ev = (uint32_t*)(wbase + L.w_ev); w.insb = wbase + L.w_insb;
	w.k0 = (uint32_t)P.seed; w.k1 = (uint32_t)(P.seed >> 32);
	w.lane = lane;
	const QualTabs qt = w.q;

	unsigned long long accBases = 0, accReads = 0, accPairs = 0, accHap = 0;
	const uint32_t insT = t.insT, delT = t.delT;
	const bool insEn = t.insEnable != 0, delEn = t.delEnable != 0;
	const uint32_t invRL = (RL > 1) ? (0xffffffffu / (uint32_t)RL + 1u) : 0xffffffffu;
	const uint32_t baseChars = t.baseChars;
	const uint32_t minQ = (uint32_t)t.minQ;

	// consecutive pairs go to consecutive warps: group g = it * gridDim + blockIdx holds FG_WORKERS pairs
	for (int group = (int)blockIdx.x; group < P.nTiles; group += (int)gridDim.x) {
		const int64_t slot = (int64_t)group * FG_WORKERS + warp;
		const int64_t e = P.emitLo + slot;
		if (e >= P.emitHi) break;
		// ---- bin of this pair (at most FG_WORKERS bins after the group's first bin)
		const int sb = P.tileStartBin[group];
		const int64_t probe = (int64_t)sb + lane + 1;
		const int64_t eb = (probe <= P.nBins) ? P.emitBase[probe] : 0x7fffffffffffffffLL;
		const int b = sb + __popc(__ballot_sync(0xffffffffu, eb <= e));
		const DevBin bin = P.bins[b];
		const int ord = (int)(e - bin.emit_base);
		const uint64_t pair = (uint64_t)(bin.plan_base + ord);
		const uint32_t fragCount = (uint32_t)(bin.frag_base + ord + 1);
		uint32_t attempt = 0;
		if (bin.risky_base >= 0) attempt = P.riskyAttempt[bin.risky_base + ord];
		w.c0 = (uint32_t)pair; w.c1 = (uint32_t)(pair >> 32);
		// ---- fragment (Segment.cpp:743-751)
		const u32x4 fb = philox_fast(w.c0, w.c1, (uint32_t)STREAM_FRAG << 24, attempt, w.k0, w.k1);
		const long long pos = f_draw_pos(fb.x, bin.spos, bin.epos);
		long long want;
		if (!t.paired) want = (long long)bin.epos - bin.spos + 1;
		else if (t.nIsize > 0) want = t.minIS + coop_lookup(s_isizeT, s_isizeSym, t.nIsize, fb.y, lane);
		else want = t.fixedInsert;
		const int64_t fstart = bin.hap_base + pos;
		const long long avail = bin.contig_end - fstart;
		const int flen = (int)(want < avail ? want : avail);
		const uint32_t posmod = (uint32_t)pos % bin.segsize;
		const bool seReverse = (!t.paired) && ((fb.z >> 31) != 0);                   // randomInteger(0, 2) != 0
		accPairs += 1;
		accHap += (unsigned long long)((flen + 3) / 4 + (flen + 7) / 8);

		// ---- prefetch the packed windows of both mates (data words lanes 0..15, mask words lanes 16..24)
		int64_t g0m[2];
		g0m[0] = seReverse ? (fstart + flen - RL) : fstart;
		g0m[1] = fstart + flen - RL;
		uint32_t wv[2];
#pragma unroll
		for (int mt = 0; mt < 2; mt++) {
			const int64_t gb = g0m[mt] - 32;
			wv[mt] = lane < 16 ? __ldg(P.hap2 + (gb >> 4) + lane) : (lane < 25 ? __ldg(P.hapN + (gb >> 5) + (lane - 16)) : 0u);
		}

		// ---- header digits: lanes 0..9 digit d of posmod, lanes 10..19 digit d of fragCount
		const uint32_t dsrc = lane < 10 ? posmod : fragCount;
		const int dpos = lane < 10 ? lane : (lane < 20 ? lane - 10 : 0);
		const uint32_t p10 = c_pow10[dpos];
		const uint32_t geMask = __ballot_sync(0xffffffffu, dsrc >= p10);          // digit d exists iff value >= 10^d
		const int nd1 = 1 + __popc(geMask & 0x3feu), nd2 = 1 + __popc(geMask & 0xff800u);
		const uint32_t dg = '0' + (dsrc / p10) % 10u;
		const int nameLen = bin.name_len;
		const int H = nameLen + nd1 + 1 + nd2 + (t.paired ? 2 : 0) + 1;
		uint32_t hb[3] = {0, 0, 0};
		const int hWords = (H + 31) >> 5;
#pragma unroll
		for (int r = 0; r < 3; r++) {
			if (r >= hWords) break;
			const int i = lane + 32 * r;
			uint32_t ch = '\n';
			int srcLane = 0;
			if (i < nameLen) ch = (uint8_t)P.names[bin.name_off + i];
			else {
				const int k = i - nameLen;
				if (k < nd1) { srcLane = nd1 - 1 - k; ch = 0; }
				else if (k == nd1) ch = '#';
				else {
					const int k2 = k - nd1 - 1;
					if (k2 < nd2) { srcLane = 10 + nd2 - 1 - k2; ch = 0; }
					else if (t.paired && k2 == nd2) ch = '/';
				}
			}
			const uint32_t dv = __shfl_sync(0xffffffffu, dg, srcLane);
			hb[r] = ch ? ch : dv;
		}

		uint32_t lens = 0;
#pragma unroll 1
		for (int mate = 0; mate < nMates; mate++) {
			const bool rev = (mate == 1) || seReverse;
			const int64_t g0 = g0m[mate];
			const int dOff = (int)((g0 - 32) & 15) + 32;   // index of base g0 relative to data word 0
			const int mOff = (int)((g0 - 32) & 31) + 32;   // same, relative to mask word 0 (32 bases per word)
			w.mDelta = mOff - dOff;
			const uint32_t c2cyc = ((uint32_t)mate << 28) | ((uint32_t)STREAM_CYCLE << 24);
			uint8_t* stage = (mate == 0 ? P.out1 : P.out2) + slot * FG_SLOT;   // this record's slot in HBM
			const uint4* subM = s_sub + ((mate == 1 && t.useCdf2) ? t.nSub : 0);
			w.sub = subM;

			// ---- phase A: one Philox block per cycle (chunks interleaved); indel tests at reference position j
			uint32_t x0[NCH], x1[NCH], x2[NCH], x3[NCH];
			philox_chunks<NCH>(w.c0, w.c1, c2cyc, (uint32_t)lane, w.k0, w.k1, x0, x1, x2, x3);
			uint32_t evbits = 0;
#pragma unroll
			for (int c = 0; c < NCH; c++) {
				const bool ins = insEn && x0[c] <= insT;                        // p <= insertRate, Profile.cpp:1560-1561
				const bool del = !ins && delEn && x1[c] <= delT;                // p2 < delRate/(1-insertRate), :1569-1570
				uint32_t hit = (ins ? 1u : 0u) | (del ? 2u : 0u);
				if (c == NCH - 1 && c * 32 + lane >= RL) hit = 0;
				evbits |= hit << (2 * c);
			}
			s_win[lane] = mate == 0 ? wv[0] : wv[1];
			__syncwarp();
			const bool slow = __any_sync(0xffffffffu, evbits != 0);

			// ---- header
#pragma unroll
			for (int r = 0; r < 3; r++) {
				if (r >= hWords) break;
				const int i = lane + 32 * r;
				if (i < H) stage[i] = (uint8_t)((t.paired && i == H - 2) ? ('1' + mate) : hb[r]);
			}
			int m = RL;
			if (!slow) {
				// ---- phase C, fast path (branch free): context straight from the packed window.
				// forward: window of cycle j starts at base g0 + j - 2; reverse: at base g0 + RL-1 - j
				const int rel0 = rev ? (dOff + RL - 1 - lane) : (dOff + lane - 2);
				const int relM0 = rev ? (mOff + RL - 1 - lane) : (mOff + lane - 2);
				const uint32_t dsh = (uint32_t)(rel0 & 15) * 2u, msh = (uint32_t)(relM0 & 31);
				const uint32_t* dptr = w.win + (rel0 >> 4);
				const uint32_t* mptr = w.win + 16 + (relM0 >> 5);
				const int dstep = rev ? -2 : 2, mstep = rev ? -1 : 1;
				const uint16_t* lut = s_lut + (rev ? 64 : 0);
				const uint32_t curBit = rev ? 1u : 4u;
				// cycles 0 and 1 have the 'X' padded contexts (Profile.cpp:1661-1666): valid context bits
				const uint32_t nmask0 = lane == 0 ? curBit : (lane == 1 ? (rev ? 3u : 6u) : 7u);
				uint8_t* st1 = stage + H + lane;
				uint8_t* st2 = stage + H + RL + 3 + lane;
				uint32_t jB = (uint32_t)(lane * B);
#pragma unroll
				for (int c = 0; c < NCH; c++) {
					const uint32_t v6 = __funnelshift_r(dptr[c * dstep], dptr[c * dstep + 1], dsh) & 63u;
					const uint32_t n3 = __funnelshift_r(mptr[c * mstep], mptr[c * mstep + 1], msh) & (c == 0 ? nmask0 : 7u);
					const uint32_t le = lut[v6];
					uint32_t row = le & 0xffu;
					const uint32_t cur = (le >> 8) & 3u;
					if (c == 0) {
						const uint32_t p1 = (le >> 10) & 3u;
						row = lane == 0 ? cur : (lane == 1 ? 4u + 4u * p1 + cur : row);
					}
					// lanes past the read end (last chunk only) are clamped so that every table index stays valid
					const uint32_t binIdx = __umulhi(c == NCH - 1 ? min(jB, (uint32_t)((RL - 1) * B)) : jB, invRL);
					jB += 32u * (uint32_t)B;
					const uint4 sr = subM[row * (uint32_t)B + binIdx];
					uint32_t call = sr.w;
					add_gt(call, x2[c], sr.x); add_gt(call, x2[c], sr.y); add_gt(call, x2[c], sr.z);
					call = n3 ? cur : call;                                        // unknown context: base passes through
					uint32_t q = qual_lookup<QP>(qt, cur, call, binIdx, B, x3[c]);
					uint32_t ch = __byte_perm(baseChars, 0, 0x4440u | call);
					if (n3 & curBit) { ch = 'N'; q = minQ + __umulhi(20u, x3[c]); }   // randomInteger(33, 53), Profile.cpp:1583
					if (c < NCH - 1 || c * 32 + lane < RL) {
						st1[c * 32] = (uint8_t)ch;
						st2[c * 32] = (uint8_t)q;
					}
				}
			} else {
#pragma unroll
				for (int c = 0; c < NCH; c++) { s_xsave[c * 32 + lane] = x2[c]; s_xsave[160 + c * 32 + lane] = x3[c]; }
				__syncwarp();
				const int relFirst = rev ? (dOff + RL - 1) : dOff;
				m = slow_read<QP>(w, evbits, mate, rev, relFirst, stage, H, &P.result->errorFlags, s_xsave, NCH * 32);
			}
			if (lane == 0) { stage[H + m] = '\n'; stage[H + m + 1] = '+'; stage[H + m + 2] = '\n'; stage[H + 2 * m + 3] = '\n'; }
			lens |= (uint32_t)(H + 2 * m + 4) << (16 * mate);
			accBases += (unsigned long long)m;
			accReads += 1;
			__syncwarp();
		}
		if (lane == 0) P.slotLens[slot] = lens;
	}

	if (lane == 0) {
		atomicAdd(&P.result->bases, accBases); atomicAdd(&P.result->reads, accReads);
		atomicAdd(&P.result->pairs, accPairs); atomicAdd(&P.result->hapBytes, accHap);
	}
}

// ---------------------------------------------------------------------------------------------
// pass 2: scan of the record lengths + copy to the exact byte offsets (dense, ordered slab)
// ---------------------------------------------------------------------------------------------
static constexpr int CP_THREADS = 256;          // pairs per tile

// copy len bytes from a 16-byte aligned global source to an arbitrarily aligned global destination
__device__ __forceinline__ void copy_realign(const uint8_t* __restrict__ src, int len, uint8_t* __restrict__ dst, int lane) {
	int head = (int)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
	if (head > len) head = len;
	if (lane < head) dst[lane] = src[lane];
	const int nvec = (len - head) >> 4;
	const uint32_t* s32 = (const uint32_t*)src;
	const int r8 = (head & 3) * 8;
	const int q0 = head >> 2;
	uint4* dv = (uint4*)(dst + head);
	for (int v = lane; v < nvec; v += 32) {
		const int q = q0 + 4 * v;
		const uint32_t w0 = s32[q], w1 = s32[q + 1], w2 = s32[q + 2], w3 = s32[q + 3], w4 = s32[q + 4];
		uint4 o;
		o.x = __funnelshift_r(w0, w1, r8);
		o.y = __funnelshift_r(w1, w2, r8);
		o.z = __funnelshift_r(w2, w3, r8);
		o.w = __funnelshift_r(w3, w4, r8);
		dv[v] = o;
	}
	const int t0 = head + (nvec << 4);
	if (lane < len - t0) dst[t0 + lane] = src[t0 + lane];
}

__global__ void __launch_bounds__(CP_THREADS) compact_kernel(const GenParams P, int nSlots, int nTiles) {
	__shared__ unsigned long long s_off[CP_THREADS];
	__shared__ unsigned long long s_warpTot[CP_THREADS / 32];
	__shared__ unsigned long long s_base;
	__shared__ int s_tile;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	volatile unsigned long long* st = P.tileState;
	while (true) {
		__syncthreads();
		if (threadIdx.x == 0) s_tile = (int)atomicAdd(P.ticket, 1u);
		__syncthreads();
		const int tile = s_tile;
		if (tile >= nTiles) break;
		const int slot = tile * CP_THREADS + threadIdx.x;
		const uint32_t lens = slot < nSlots ? P.slotLens[slot] : 0u;
		const unsigned long long v = ((unsigned long long)(lens & 0xffffu) << 32) | (lens >> 16);
		unsigned long long incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += o;
		}
		if (lane == 31) s_warpTot[warp] = incl;
		__syncthreads();
		unsigned long long wbase = 0, tot = 0;
#pragma unroll
		for (int i = 0; i < CP_THREADS / 32; i++) { if (i < warp) wbase += s_warpTot[i]; tot += s_warpTot[i]; }
		s_off[threadIdx.x] = wbase + incl - v;
		if (warp == 0) {
			const unsigned long long packed = ((tot >> 32) << 31) | (tot & 0x7fffffffull);
			unsigned long long excl = 0;
			if (tile == 0) {
				if (lane == 0) st[0] = ST_P | packed;
			} else {
				if (lane == 0) st[tile] = ST_A | packed;
				int pred = tile - 1;
				while (true) {
					const int idx = pred - lane;
					unsigned long long sv = idx >= 0 ? st[idx] : ST_P;
					while (__any_sync(0xffffffffu, (sv & ST_MASK) == 0ull)) sv = idx >= 0 ? st[idx] : ST_P;
					const unsigned pm = __ballot_sync(0xffffffffu, (sv & ST_MASK) == ST_P);
					const int firstP = pm ? (__ffs(pm) - 1) : 32;
					unsigned long long contrib = (lane <= firstP) ? (sv & ~ST_MASK) : 0ull;
#pragma unroll
					for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
					excl += contrib;
					if (pm) break;
					pred -= 32;
				}
				if (lane == 0) st[tile] = ST_P | (excl + packed);
			}
			if (lane == 0) {
				s_base = excl;
				if (tile == nTiles - 1) {
					const unsigned long long fin = excl + packed;
					P.result->bytes1 = fin >> 31;
					P.result->bytes2 = fin & 0x7fffffffull;
					if ((fin >> 31) > P.cap1 || (fin & 0x7fffffffull) > P.cap2) atomicOr(&P.result->errorFlags, 1u);
				}
			}
		}
		__syncthreads();
		const unsigned long long tb = s_base;
		const unsigned long long g1b = tb >> 31, g2b = tb & 0x7fffffffull;
		for (int r = 0; r < 32; r++) {
			const int i = warp * 32 + r;
			const int sl = tile * CP_THREADS + i;
			if (sl >= nSlots) break;
			const uint32_t ln = __shfl_sync(0xffffffffu, lens, r);
			const unsigned long long off = s_off[i];
			const int l1 = (int)(ln & 0xffffu), l2 = (int)(ln >> 16);
			const unsigned long long d1 = g1b + (off >> 32), d2 = g2b + (off & 0xffffffffull);
			if (d1 + (unsigned)l1 <= P.cap1 && d2 + (unsigned)l2 <= P.cap2) {
				copy_realign(P.out1 + (size_t)sl * FG_SLOT, l1, P.dense1 + d1, lane);
				if (l2) copy_realign(P.out2 + (size_t)sl * FG_SLOT, l2, P.dense2 + d2, lane);
			}
		}
	}
}

bool fast_supported(const DevTables& t, int smemLimit, int* qmode, size_t* smemBytes) {
	if (t.N != 4 || t.K != 3 || t.RL > 160 || t.RL < 33 || t.nIsize > 1024 || t.B > 160) return false;
	const int nSubTotal = t.nSub * (t.useCdf2 ? 2 : 1);
	const int full = fast_layout(nSubTotal, t.nQualRows * 8, t.nIsize, t.nInsLen, t.nDelLen).total;
	const int diag = fast_layout(nSubTotal, 4 * t.B * t.qualDiagPitch, t.nIsize, t.nInsLen, t.nDelLen).total;
	const int none = fast_layout(nSubTotal, 0, t.nIsize, t.nInsLen, t.nDelLen).total;
	if (t.qualPitch == 8 && full <= smemLimit) { *qmode = 8; *smemBytes = (size_t)full; return true; }
	if (diag <= smemLimit) { *qmode = 2; *smemBytes = (size_t)diag; return true; }
	if (none <= smemLimit) { *qmode = 0; *smemBytes = (size_t)none; return true; }
	return false;
}

template <int NCH, int QP>
static cudaError_t launch_fast_variant(const GenParams& P, size_t smemBytes, int grid, cudaStream_t stream) {
	auto kern = generate_slots_kernel<NCH, QP>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
	if (e != cudaSuccess) return e;
	kern<<<grid, FG_THREADS, smemBytes, stream>>>(P);
	return cudaGetLastError();
}

// P.out1/out2 = slot scratch, P.dense1/dense2 = final slabs, P.nTiles = groups of FG_WORKERS pairs
cudaError_t launch_generate_fast(const GenParams& P, int qmode, size_t smemBytes, int grid, int smCount, cudaStream_t stream,
                                 cudaEvent_t e0, cudaEvent_t e1, cudaEvent_t e2) {
	const int nch = (P.t.RL + 31) / 32;
	cudaError_t e;
	if (e0) cudaEventRecord(e0, stream);
	if (qmode == 8) {
		if (nch <= 3) e = launch_fast_variant<3, 8>(P, smemBytes, grid, stream);
		else if (nch == 4) e = launch_fast_variant<4, 8>(P, smemBytes, grid, stream);
		else e = launch_fast_variant<5, 8>(P, smemBytes, grid, stream);
	} else if (qmode == 2) {
		if (nch <= 3) e = launch_fast_variant<3, 2>(P, smemBytes, grid, stream);
		else if (nch == 4) e = launch_fast_variant<4, 2>(P, smemBytes, grid, stream);
		else e = launch_fast_variant<5, 2>(P, smemBytes, grid, stream);
	} else {
		if (nch <= 3) e = launch_fast_variant<3, 0>(P, smemBytes, grid, stream);
		else if (nch == 4) e = launch_fast_variant<4, 0>(P, smemBytes, grid, stream);
		else e = launch_fast_variant<5, 0>(P, smemBytes, grid, stream);
	}
	if (e != cudaSuccess) return e;
	if (e1) cudaEventRecord(e1, stream);
	const int nSlots = (int)(P.emitHi - P.emitLo);
	const int nTiles = (nSlots + CP_THREADS - 1) / CP_THREADS;
	int cgrid = smCount * 8;
	if (cgrid > nTiles) cgrid = nTiles;
	compact_kernel<<<cgrid, CP_THREADS, 0, stream>>>(P, nSlots, nTiles);
	if (e2) cudaEventRecord(e2, stream);
	return cudaGetLastError();
}

}  // namespace ssc
