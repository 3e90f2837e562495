// api.cu -- C ABI of include/simuscop.h: device memory, plan upload, census, batched
// generation with pinned double-buffered device->host streaming.  No CPU fallback: every
// entry point needs a CUDA device and fails loudly otherwise.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/simuscop.h"
#include "device_types.h"
#include "kernels.h"
#include "tables.h"
#include "gz.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if (e__ != cudaSuccess)                                                                    \
			return fail(SSC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

template <class T> struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	cudaError_t alloc(size_t count) {
		release();
		n = count;
		if (count == 0) return cudaSuccess;
		return cudaMalloc((void**)&p, count * sizeof(T));
	}
	cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
		cudaError_t e = alloc(v.size());
		if (e != cudaSuccess || v.empty()) return e;
		return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct ssc_handle {
	int device = 0;
	int smCount = 0;
	int smemLimit = 0;
	cudaStream_t compute = nullptr, copy = nullptr, copy2 = nullptr;   // copy/copy2: device->host of file 1 / file 2 (two copy engines)
	cudaEvent_t evStart = nullptr, evStop = nullptr, evGen[2] = {nullptr, nullptr}, evCopy[2] = {nullptr, nullptr}, evCopy2[2] = {nullptr, nullptr};

	// options
	int64_t batchPairs = 1 << 20;
	bool fp64 = false;
	bool forceGeneric = false;
	bool noSplice = false;        // tests: see GenParams::noSplice
	int maxCtas = 0;              // > 0: cap the grid of the generation kernel (tests: many tickets per warp on small inputs)
	int moverCtas = -1;           // "concurrent_move": pass 2b of batch k runs on that many SMs (bulk-copy mover, second stream) under the
	                              // generation kernel of batch k+1, which leaves them free.  -1 = default: 8 of 148 SMs (measured
	                              // optimum on the 3 Gb job: 4.30 -> 3.97 ms per step; 6 SMs cannot keep up, 12 cost the generation more)
	cudaStream_t mover = nullptr;
	cudaEvent_t evPre = nullptr, evScanned[2] = {nullptr, nullptr}, evMoved[2] = {nullptr, nullptr};
	bool noQ16 = false;           // "no_q16": keep profiles with 9..40 live quality symbols on the diagonal-rows mode (A/B, tests)
	bool prefetchWindows = true;  // the ticket prologue of the fast kernel prefetches its pairs' haplotype windows into the L2
	bool carryPass2 = false;      // on: pass 2b (blob moves) of batch k rides on the generation kernel of batch k+1 instead of a stand-alone
	                              // kernel per batch.  Measured neutral on the 3 Gb job: a warp of the generation kernel
	                              // moves its 21 KB at the latency-bound rate of one warp, which costs what the stand-alone kernel costs
	bool gzip = false;            // slabs hold gzip members (one per ticket blob) instead of plain FASTQ
	bool haveGz = false;          // Huffman / CRC tables of the current plan are on the device
	ssc::GzTables* d_gzTab = nullptr;
	uint8_t* d_gzBlobs[2] = {nullptr, nullptr};
	DevBuf<unsigned long long> d_gzLens;

	// profile
	bool haveProfile = false;
	ssc::DeviceTablesHost th;
	ssc::DevTables dt;
	DevBuf<uint32_t> d_isizeT, d_insT, d_delT, d_qualT, d_qualDiagT;
	DevBuf<uint8_t> d_qualDiagSym;
	DevBuf<uint16_t> d_isizeSym, d_insSym, d_delSym;
	DevBuf<uint8_t> d_qualSym;
	DevBuf<uint4> d_sub;
	DevBuf<double> d_fIsize, d_fIns, d_fDel, d_fSub1, d_fSub2, d_fQual;
	DevBuf<int8_t> d_lut;

	// genome
	DevBuf<uint32_t> d_hap2, d_hapN;
	uint64_t genomeCap = 0, genomeSize = 0;
	DevBuf<uint8_t> d_ref;                        // ASCII chromosome for ssc_genome_append_ref
	DevBuf<uint8_t> d_raw;                        // raw FASTA lines of the chromosome (ssc_reference_upload_fasta), grow-only
	DevBuf<unsigned long long> d_other;
	// ssc_reference_prefetch_fasta: the next record is read and unfolded into a second set of buffers by a background thread
	// on its own stream while the caller works with the current reference; ssc_reference_adopt_prefetched swaps the sets
	DevBuf<uint8_t> d_ref2, d_raw2;
	DevBuf<unsigned long long> d_other2;
	cudaStream_t upload = nullptr;
	cudaEvent_t evRefFree = nullptr;
	std::thread prefetchThread;
	bool prefetching = false;
	int prefetchRc = 0;
	std::string prefetchErr;
	uint64_t prefetchBases = 0, prefetchRaw = 0, prefetchOther = 0;
	uint64_t refSize = 0;
	uint8_t* h_stage[2] = {nullptr, nullptr};   // pinned upload staging
	uint8_t* d_stage[2] = {nullptr, nullptr};
	cudaEvent_t evStage[2] = {nullptr, nullptr};
	static constexpr int FA_READERS = 2;          // ssc_reference_upload_fasta: reader threads, one pinned buffer each (four readers
	                                              // of 16 MB chunks were slower on the 16-vCPU boxes: 0.69 vs 0.28-0.31 s for the 3 Gb FASTA)
	uint8_t* h_fa[2 * FA_READERS] = {nullptr, nullptr, nullptr, nullptr};       // [0, FA_READERS): foreground, the rest: prefetch
	cudaEvent_t evFa[2 * FA_READERS] = {nullptr, nullptr, nullptr, nullptr};
	size_t faBytes = 32u << 20;
	size_t stageBytes = 32u << 20;

	// plan
	bool havePlan = false;
	uint64_t seed = 0;
	std::vector<int64_t> planBaseAll, emitBaseAll;   // per original bin (+1 sentinel)
	std::vector<int64_t> devEmitBase;                // per device bin (+1 sentinel): host copy of d_emitBase
	DevBuf<ssc::DevBin> d_bins;
	DevBuf<int64_t> d_emitBase;
	DevBuf<uint16_t> d_risky;
	DevBuf<char> d_names;
	int64_t nDevBins = 0, plannedPairs = 0, emittedPairs = 0;
	int maxRecBytes = 0;

	// batch resources
	int64_t slabPairs = 0;
	uint64_t slabCap = 0;
	uint8_t* d_out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][file]
	uint8_t* h_out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
	DevBuf<int32_t> d_tileStart[2];
	DevBuf<unsigned long long> d_tileState[2];
	DevBuf<unsigned int> d_ticket[2];
	uint8_t* d_slots[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][file] pass-1 blob scratch of the fast kernel; file 2 lies
	                                                                     // behind file 1 in the same allocation.  Two buffers: the blobs of batch
	                                                                     // k are moved to their dense slab by the launch that generates batch k+1
	DevBuf<unsigned int> d_ticket2;
	DevBuf<unsigned long long> d_blobPrefix[2];
	struct { bool valid = false; ssc::GenParams P; } pending;          // generated + scanned, blobs not yet moved
	cudaEvent_t evDense[2] = {nullptr, nullptr};
	DevBuf<int64_t> d_cenStarts;                  // ssc_gc_census scratch (grow-only)
	DevBuf<int32_t> d_cenLens, d_cenGc, d_cenNn;
	ssc::BatchResult* d_result[2] = {nullptr, nullptr};
	ssc::BatchResult* h_result[2] = {nullptr, nullptr};

	std::vector<cudaEvent_t> kev;   // per-batch kernel timing events of ssc_generate_device (3 per batch)
	int kevUsed = 0;
	bool timeKernels = false;

	ssc_stats stats;
};

namespace {

bool use_fast(const ssc_handle* h, int* qmode, size_t* smemBytes) {
	if (h->fp64 || h->forceGeneric) return false;
	return ssc::fast_supported(h->dt, h->smemLimit, qmode, smemBytes);
}


int ensure_batch_resources(ssc_handle* h, bool needHost) {
	int64_t pairs = std::min<int64_t>(h->batchPairs, std::max<int64_t>(h->emittedPairs, 1));
	pairs = ((pairs + GEN_TILE_PAIRS - 1) / GEN_TILE_PAIRS) * GEN_TILE_PAIRS;   // multiple of both tile sizes
	uint64_t cap = (uint64_t)pairs * (uint64_t)h->maxRecBytes + 4096;
	if (cap >= (1ull << 31)) {
		pairs = (int64_t)(((1ull << 31) - 8192) / (uint64_t)h->maxRecBytes);
		pairs = (pairs / GEN_TILE_PAIRS) * GEN_TILE_PAIRS;
		if (pairs <= 0) return fail(SSC_ERR_INVALID, "record size too large");
		cap = (uint64_t)pairs * (uint64_t)h->maxRecBytes + 4096;
	}
	int nFiles = h->dt.paired ? 2 : 1;
	{
		// the fast kernel addresses its blob scratch (FG_SLOT bytes per record, both files) with 32-bit offsets
		const int64_t maxFast = (int64_t)(((1ull << 32) - (1ull << 20)) / ((uint64_t)FG_SLOT * (uint64_t)nFiles));
		if (pairs > maxFast) {
			pairs = (maxFast / GEN_TILE_PAIRS) * GEN_TILE_PAIRS;
			cap = (uint64_t)pairs * (uint64_t)h->maxRecBytes + 4096;
		}
	}
	if (pairs != h->slabPairs || cap != h->slabCap) {
		for (int b = 0; b < 2; b++)
			for (int f = 0; f < 2; f++) {
				if (h->d_out[b][f]) cudaFree(h->d_out[b][f]);
				if (h->h_out[b][f]) cudaFreeHost(h->h_out[b][f]);
				h->d_out[b][f] = nullptr; h->h_out[b][f] = nullptr;
			}
		for (int b = 0; b < 2; b++) { if (h->d_slots[b][0]) cudaFree(h->d_slots[b][0]); h->d_slots[b][0] = h->d_slots[b][1] = nullptr; }
		h->slabPairs = 0; h->slabCap = 0;          // committed below, once every allocation has succeeded
		h->pending.valid = false;
		for (int b = 0; b < 2; b++) {
			// file 2's blobs behind file 1's: the kernel addresses both from one base pointer with 32-bit cursors
			CK(cudaMalloc((void**)&h->d_slots[b][0], ((size_t)pairs * FG_SLOT + 256) * nFiles));
			h->d_slots[b][1] = nFiles == 2 ? h->d_slots[b][0] + ((size_t)pairs * FG_SLOT + 256) : nullptr;
			CK(h->d_blobPrefix[b].alloc((size_t)(pairs / 16) + 2));
		}
		for (int f = 0; f < 2; f++) { if (h->d_gzBlobs[f]) cudaFree(h->d_gzBlobs[f]); h->d_gzBlobs[f] = nullptr; }
		CK(h->d_ticket2.alloc(1));
		int nTiles = (int)(pairs / 16) + 2;   // enough for every kernel's tile size
		for (int b = 0; b < 2; b++) {
			for (int f = 0; f < nFiles; f++) CK(cudaMalloc((void**)&h->d_out[b][f], cap));
			CK(h->d_tileStart[b].alloc(nTiles));
			CK(h->d_tileState[b].alloc(nTiles));
			CK(h->d_ticket[b].alloc(1));
		}
		h->slabPairs = pairs; h->slabCap = cap;
	}
	if (h->gzip && !h->d_gzBlobs[0]) {
		for (int f = 0; f < 2; f++) CK(cudaMalloc((void**)&h->d_gzBlobs[f], (size_t)h->slabPairs * FG_SLOT + 64));
		CK(h->d_gzLens.alloc((size_t)(h->slabPairs / 16) + 2));
		if (!h->d_gzTab) CK(cudaMalloc((void**)&h->d_gzTab, sizeof(ssc::GzTables)));
	}
	if (needHost)
		for (int b = 0; b < 2; b++)
			for (int f = 0; f < nFiles; f++)
				if (!h->h_out[b][f]) CK(cudaMallocHost((void**)&h->h_out[b][f], h->slabCap));
	return SSC_OK;
}

// emitted-pair index of the first emitted pair whose planned ID is >= p
int64_t emit_index_of_plan(const ssc_handle* h, int64_t p) {
	if (p <= 0) return 0;
	if (p >= h->plannedPairs) return h->emittedPairs;
	const std::vector<int64_t>& pb = h->planBaseAll;
	size_t b = std::upper_bound(pb.begin(), pb.end(), p) - pb.begin() - 1;   // pb[b] <= p < pb[b+1]
	int64_t emitCount = h->emitBaseAll[b + 1] - h->emitBaseAll[b];
	return h->emitBaseAll[b] + std::min<int64_t>(p - pb[b], emitCount);
}

// Fits the gzip Huffman table of a plan to the first tickets of its first batch (the blobs are in scratch already).
int build_gz_tables(ssc_handle* h, const ssc::GenParams& P, int nTiles) {
	const int sample = std::min(nTiles, 256);
	const size_t pitch = P.blobPitch;
	std::vector<unsigned long long> lens((size_t)sample);
	std::vector<uint8_t> blob((size_t)sample * pitch);
	uint64_t hist[256];
	memset(hist, 0, sizeof(hist));
	CK(cudaStreamSynchronize(h->compute));
	CK(cudaMemcpy(lens.data(), P.tileState, lens.size() * 8, cudaMemcpyDeviceToHost));
	for (int f = 0; f < (h->dt.paired ? 2 : 1); f++) {
		CK(cudaMemcpy(blob.data(), f ? P.out2 : P.out1, blob.size(), cudaMemcpyDeviceToHost));
		for (int j = 0; j < sample; j++) {
			const size_t n = f ? (size_t)(lens[j] & 0x7fffffffull) : (size_t)(lens[j] >> 31);
			const uint8_t* p = blob.data() + (size_t)j * pitch;
			for (size_t i = 0; i < n; i++) hist[p[i]]++;
		}
	}
	ssc::GzTables tab;
	const char* err = ssc::gz_build_tables(hist, &tab);
	if (err[0]) return fail(SSC_ERR_INVALID, "gzip tables: %s", err);
	CK(cudaMemcpyAsync(h->d_gzTab, &tab, sizeof(tab), cudaMemcpyHostToDevice, h->compute));   // stream order with the deflate kernel
	CK(cudaStreamSynchronize(h->compute));          // tab lives on this stack frame
	h->haveGz = true;
	return SSC_OK;
}

// Moves the blobs of the last launched batch to its dense slab with the stand-alone kernel (the last batch of a call: no
// later launch carries its moves).
int flush_pending(ssc_handle* h) {
	if (!h->pending.valid) return SSC_OK;
	h->pending.valid = false;
	CK(ssc::launch_move_blobs(h->pending.P, h->smCount, h->compute));
	h->stats.launches += 1;
	return SSC_OK;
}

// Launches batch [emitLo, emitHi) into buffer set `buf`.  Fast kernel, plain output: pass 1 + the scan of the blob lengths;
// the blobs are moved to the dense slab d_out[buf] by the next launch_batch (fused into its generation kernel) or by
// flush_pending.  Every other mode leaves the finished slab in d_out[buf] when its kernels have run.
int launch_batch(ssc_handle* h, int buf, int64_t emitLo, int64_t emitHi) {
	int qsmem = 0; size_t fastSmem = 0;
	const bool fast = use_fast(h, &qsmem, &fastSmem);
	const int tp = fast ? FG_CHUNK : GEN_TILE_PAIRS;
	int nTiles = (int)((emitHi - emitLo + tp - 1) / tp);
	cudaStream_t s = h->compute;
	CK(cudaMemsetAsync(h->d_tileState[buf].p, 0, sizeof(unsigned long long) * nTiles, s));
	CK(cudaMemsetAsync(h->d_ticket[buf].p, 0, sizeof(unsigned int), s));
	CK(cudaMemsetAsync(h->d_ticket2.p, 0, sizeof(unsigned int), s));
	CK(cudaMemsetAsync(h->d_result[buf], 0, sizeof(ssc::BatchResult), s));
	CK(ssc::launch_locate(h->d_emitBase.p, h->nDevBins, emitLo, tp, nTiles, h->d_tileStart[buf].p, s));
	h->stats.launches += 1;
	{
		// distinct bins of the batch (statistics: the algorithmic bin-record bytes of the roofline)
		const std::vector<int64_t>& eb = h->devEmitBase;
		const size_t bLo = std::upper_bound(eb.begin(), eb.end(), emitLo) - eb.begin() - 1;
		const size_t bHi = std::lower_bound(eb.begin(), eb.end(), emitHi) - eb.begin();
		h->stats.bin_bytes += (uint64_t)(bHi > bLo ? bHi - bLo : 0) * sizeof(ssc::DevBin);
	}
	ssc::GenParams P;
	memset(&P, 0, sizeof(P));
	P.t = h->dt;
	P.hap2 = h->d_hap2.p; P.hapN = h->d_hapN.p;
	P.bins = h->d_bins.p; P.emitBase = h->d_emitBase.p; P.nBins = h->nDevBins;
	P.riskyAttempt = h->d_risky.p; P.names = h->d_names.p;
	P.seed = h->seed; P.emitLo = emitLo; P.emitHi = emitHi; P.one = 1; P.noSplice = h->noSplice ? 1 : 0;
	P.qstride = (fast && qsmem == 8) ? (uint32_t)h->dt.qualBins * 68u : 68u;   // F_QROW of gen_fast.cu (unused by the 16-bit-key mode)
	P.insLim = h->dt.insEnable ? h->dt.insT + 1u : 0u;
	P.delLim = h->dt.delEnable ? h->dt.delT + 1u : 0u;
	P.alwaysSlow = (h->dt.insEnable && h->dt.insT == 0xFFFFFFFFu) || (h->dt.delEnable && h->dt.delT == 0xFFFFFFFFu);
	for (int r = 0; r < 10; r++) {
		P.rk[2 * r] = (uint32_t)h->seed + (uint32_t)r * 0x9E3779B9u;
		P.rk[2 * r + 1] = (uint32_t)(h->seed >> 32) + (uint32_t)r * 0xBB67AE85u;
	}
	P.tileStartBin = h->d_tileStart[buf].p; P.nTiles = nTiles; P.nLoop = nTiles; P.nTilesPrev = 0;
	P.prefetchWindows = h->prefetchWindows ? 1 : 0;
	P.tileState = h->d_tileState[buf].p; P.ticket = h->d_ticket[buf].p; P.ticket2 = h->d_ticket2.p; P.blobPrefix = h->d_blobPrefix[buf].p;
	P.out1 = h->d_out[buf][0]; P.out2 = h->d_out[buf][1];
	P.cap1 = h->slabCap; P.cap2 = h->slabCap;
	P.result = h->d_result[buf];
	int grid = std::min(nTiles, h->smCount);
	if (fast) {
		// pass 1 writes one blob per ticket and file into the scratch, pass 2 (scan of the blob lengths + one move per blob) the dense slab
		P.dense1 = P.out1; P.dense2 = P.out2;
		P.out1 = h->d_slots[buf][0]; P.out2 = h->d_slots[buf][1];
		P.blobPitch = (uint32_t)(FG_CHUNK * FG_SLOT);
		P.file2Off = (uint32_t)(h->d_slots[buf][1] ? h->d_slots[buf][1] - h->d_slots[buf][0] : 0);
		cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
		if (h->timeKernels) {
			while ((int)h->kev.size() < h->kevUsed + 3) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->kev.push_back(e); }
			e0 = h->kev[h->kevUsed]; e1 = h->kev[h->kevUsed + 1]; e2 = h->kev[h->kevUsed + 2];
			h->kevUsed += 3;
		}
		if (!h->gzip) {
			int movedBuf = -1;
			if (h->pending.valid && h->moverCtas > 0 && h->moverCtas < h->smCount) {
				// pass 2b of the previous batch on the mover stream, under this batch's generation kernel (which leaves it the SMs)
				const ssc::GenParams& Q = h->pending.P;
				movedBuf = buf ^ 1;
				CK(cudaEventRecord(h->evPre, s));                          // everything this launch had to wait for (copies, ...)
				CK(cudaStreamWaitEvent(h->mover, h->evPre, 0));
				CK(cudaStreamWaitEvent(h->mover, h->evScanned[movedBuf], 0));
				CK(ssc::launch_move_blobs_tma(Q, h->moverCtas, h->mover));
				CK(cudaEventRecord(h->evMoved[movedBuf], h->mover));
				h->pending.valid = false;
				h->stats.launches += 1;
			}
			if (h->pending.valid && h->carryPass2) {
				// this launch carries pass 2b of the previous batch
				const ssc::GenParams& Q = h->pending.P;
				P.nTilesPrev = Q.nTiles; P.nLoop = std::max(nTiles, Q.nTiles);
				P.prevBlobs = Q.out1; P.prevFile2Off = Q.file2Off; P.prevBlobPitch = Q.blobPitch;
				P.prevTileState = Q.tileState; P.prevPrefix = Q.blobPrefix;
				P.prevDense1 = Q.dense1; P.prevDense2 = Q.dense2; P.prevCap1 = Q.cap1; P.prevCap2 = Q.cap2;
				h->pending.valid = false;
			}
			if (h->pending.valid) { int rc = flush_pending(h); if (rc) return rc; }     // (neither carried nor moved concurrently)
			grid = std::min((P.nLoop + FG_GEN - 1) / FG_GEN, h->smCount - (movedBuf >= 0 ? h->moverCtas : 0));
			if (h->maxCtas > 0) grid = std::min(grid, h->maxCtas);
			CK(ssc::launch_generate_fast(P, qsmem, fastSmem, grid, s, e0, e1));
			CK(ssc::launch_scan_blobs(P, s));
			CK(cudaEventRecord(h->evScanned[buf], s));
			h->pending.P = P; h->pending.valid = true;
			h->stats.launches += 2;
			if (!h->carryPass2 && h->moverCtas <= 0) { int rc = flush_pending(h); if (rc) return rc; }
			if (e2) CK(cudaEventRecord(e2, s));
			// later work on this stream (the next launch reuses the blob scratch, copies read the dense slab) comes after the mover
			if (movedBuf >= 0) CK(cudaStreamWaitEvent(s, h->evMoved[movedBuf], 0));
		} else {
			// gzip mode: blobs -> (first batch of a plan: fit the Huffman table to a sample) -> one gzip member per blob -> pass 2 on the members
			int rc = flush_pending(h);
			if (rc) return rc;
			grid = std::min((nTiles + FG_GEN - 1) / FG_GEN, h->smCount);
			if (h->maxCtas > 0) grid = std::min(grid, h->maxCtas);
			CK(ssc::launch_generate_fast(P, qsmem, fastSmem, grid, s, e0, e1));
			if (!h->haveGz) { rc = build_gz_tables(h, P, nTiles); if (rc) return rc; }
			CK(cudaMemsetAsync(h->d_gzLens.p, 0, sizeof(unsigned long long) * nTiles, s));
			CK(ssc::launch_deflate_blobs(P.out1, P.out2, P.tileState, nTiles, P.blobPitch, h->d_gzBlobs[0], h->d_gzBlobs[1], h->d_gzLens.p,
			                             FG_CHUNK * FG_SLOT, h->d_gzTab, &P.result->errorFlags, h->smCount, s));
			ssc::GenParams P2 = P;
			P2.out1 = h->d_gzBlobs[0]; P2.out2 = h->d_gzBlobs[1]; P2.tileState = h->d_gzLens.p; P2.blobPitch = FG_CHUNK * FG_SLOT;
			CK(ssc::launch_pass2(P2, h->smCount, s));
			if (e2) CK(cudaEventRecord(e2, s));
			h->stats.launches += 4;
		}
	} else {
		if (h->gzip) return fail(SSC_ERR_INVALID, "gzip output needs the fast kernel (kmer 3, read length 33..160)");
		int rc = flush_pending(h);
		if (rc) return rc;
		ssc::GenVariant v = ssc::choose_variant(h->dt, h->fp64, h->smemLimit);
		if (!v.ok) return fail(SSC_ERR_INVALID, "no kernel variant for read length %d / kmer %d", h->dt.RL, h->dt.K);
		CK(ssc::launch_generate(P, v, grid, s));
		h->stats.launches += 1;
	}
	CK(cudaMemcpyAsync(h->h_result[buf], h->d_result[buf], sizeof(ssc::BatchResult), cudaMemcpyDeviceToHost, s));
	h->stats.gen_launches += 1;
	return SSC_OK;
}

// error exit of a pipelined call: nothing of it may still be running when the caller sees the code
int bail(ssc_handle* h, int rc) {
	const std::string msg = g_err;
	cudaDeviceSynchronize();
	h->pending.valid = false;
	g_err = msg;
	return rc;
}

int check_result(ssc_handle* h, const ssc::BatchResult& r) {
	if (r.errorFlags & 1u) return fail(SSC_ERR_OVERFLOW, "output slab overflow (internal capacity estimate too small)");
	if (r.errorFlags & 2u) return fail(SSC_ERR_OVERFLOW, "a read outgrew the per-read scratch (insertions > 96 bases)");
	if (r.errorFlags & 4u) return fail(SSC_ERR_OVERFLOW, "more than 32 indel events or 128 inserted bases in one read");
	if (r.errorFlags & 8u) return fail(SSC_ERR_OVERFLOW, "a gzip member outgrew its scratch blob");
	h->stats.pairs_emitted += r.pairs;
	h->stats.reads_emitted += r.reads;
	h->stats.bases_emitted += r.bases;
	h->stats.fastq_bytes += h->gzip ? r.rawBytes : r.bytes1 + r.bytes2;
	if (h->gzip) h->stats.gz_bytes += r.bytes1 + r.bytes2;
	h->stats.hap_bytes += r.hapBytes;
	return SSC_OK;
}

}  // namespace

extern "C" {

const char* ssc_last_error(void) { return g_err.c_str(); }
int ssc_version(void) { return 1; }

static int init_handle(ssc_handle* h, int device) {
	h->device = device;
	memset(&h->stats, 0, sizeof(h->stats));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	h->smCount = prop.multiProcessorCount;
	if (h->moverCtas < 0) h->moverCtas = h->smCount >= 64 ? (h->smCount * 8 + 74) / 148 : 0;
	h->smemLimit = (int)prop.sharedMemPerBlockOptin - 1024;
	CK(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&h->copy2, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&h->mover, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&h->evPre, cudaEventDisableTiming));
	CK(cudaEventCreate(&h->evStart));
	CK(cudaEventCreate(&h->evStop));
	for (int i = 0; i < 2; i++) {
		CK(cudaEventCreateWithFlags(&h->evGen[i], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->evCopy[i], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->evCopy2[i], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->evStage[i], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->evDense[i], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->evScanned[i], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->evMoved[i], cudaEventDisableTiming));
		CK(cudaMalloc((void**)&h->d_result[i], sizeof(ssc::BatchResult)));
		CK(cudaMallocHost((void**)&h->h_result[i], sizeof(ssc::BatchResult)));
	}
	return SSC_OK;
}

int ssc_create(int device, ssc_handle** out) {
	if (!out) return fail(SSC_ERR_INVALID, "out is null");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(SSC_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
		            e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	if (device < 0 || device >= count) return fail(SSC_ERR_INVALID, "device %d out of range (%d devices)", device, count);
	CK(cudaSetDevice(device));
	ssc_handle* h = new ssc_handle();
	const int rc = init_handle(h, device);
	if (rc) { const std::string msg = g_err; ssc_destroy(h); g_err = msg; return rc; }   // nothing of a half-built handle leaks
	*out = h;
	return SSC_OK;
}

int ssc_destroy(ssc_handle* h) {
	if (!h) return SSC_OK;
	if (h->prefetchThread.joinable()) h->prefetchThread.join();
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	for (int b = 0; b < 2; b++) {
		for (int f = 0; f < 2; f++) {
			if (h->d_out[b][f]) cudaFree(h->d_out[b][f]);
			if (h->h_out[b][f]) cudaFreeHost(h->h_out[b][f]);
		}
		if (h->h_stage[b]) cudaFreeHost(h->h_stage[b]);
		if (h->d_stage[b]) cudaFree(h->d_stage[b]);
		for (int r = b; r < 2 * ssc_handle::FA_READERS; r += 2) {
			if (h->h_fa[r]) cudaFreeHost(h->h_fa[r]);
			if (h->evFa[r]) cudaEventDestroy(h->evFa[r]);
		}
		if (h->d_slots[b][0]) cudaFree(h->d_slots[b][0]);       // d_slots[b][1] points into the same allocation
		h->d_blobPrefix[b].release();
		if (h->evDense[b]) cudaEventDestroy(h->evDense[b]);
		if (h->evScanned[b]) cudaEventDestroy(h->evScanned[b]);
		if (h->evMoved[b]) cudaEventDestroy(h->evMoved[b]);
		if (h->d_result[b]) cudaFree(h->d_result[b]);
		if (h->h_result[b]) cudaFreeHost(h->h_result[b]);
		h->d_tileStart[b].release(); h->d_tileState[b].release(); h->d_ticket[b].release();
		if (h->evGen[b]) cudaEventDestroy(h->evGen[b]);
		if (h->evCopy[b]) cudaEventDestroy(h->evCopy[b]);
		if (h->evCopy2[b]) cudaEventDestroy(h->evCopy2[b]);
		if (h->evStage[b]) cudaEventDestroy(h->evStage[b]);
	}
	h->d_isizeT.release(); h->d_insT.release(); h->d_delT.release(); h->d_qualT.release();
	h->d_qualDiagT.release(); h->d_qualDiagSym.release();
	h->d_isizeSym.release(); h->d_insSym.release(); h->d_delSym.release(); h->d_qualSym.release();
	h->d_sub.release(); h->d_fIsize.release(); h->d_fIns.release(); h->d_fDel.release();
	h->d_fSub1.release(); h->d_fSub2.release(); h->d_fQual.release(); h->d_lut.release();
	h->d_hap2.release(); h->d_hapN.release(); h->d_ref.release(); h->d_raw.release(); h->d_other.release();
	h->d_ref2.release(); h->d_raw2.release(); h->d_other2.release();
	if (h->upload) cudaStreamDestroy(h->upload);
	if (h->evRefFree) cudaEventDestroy(h->evRefFree);
	h->d_ticket2.release(); h->d_gzLens.release();
	for (int f = 0; f < 2; f++) if (h->d_gzBlobs[f]) cudaFree(h->d_gzBlobs[f]);
	if (h->d_gzTab) cudaFree(h->d_gzTab);
	h->d_cenStarts.release(); h->d_cenLens.release(); h->d_cenGc.release(); h->d_cenNn.release();
	h->d_bins.release(); h->d_emitBase.release(); h->d_risky.release(); h->d_names.release();
	for (cudaEvent_t e : h->kev) cudaEventDestroy(e);
	if (h->evStart) cudaEventDestroy(h->evStart);
	if (h->evStop) cudaEventDestroy(h->evStop);
	if (h->compute) cudaStreamDestroy(h->compute);
	if (h->copy) cudaStreamDestroy(h->copy);
	if (h->copy2) cudaStreamDestroy(h->copy2);
	if (h->mover) cudaStreamDestroy(h->mover);
	if (h->evPre) cudaEventDestroy(h->evPre);
	delete h;
	return SSC_OK;
}

int ssc_set_option(ssc_handle* h, const char* key, int64_t value) {
	if (!h || !key) return fail(SSC_ERR_INVALID, "null argument");
	if (!strcmp(key, "batch_pairs")) {
		if (value < GEN_TILE_PAIRS) return fail(SSC_ERR_INVALID, "batch_pairs must be >= %d", GEN_TILE_PAIRS);
		h->batchPairs = value;
		return SSC_OK;
	}
	if (!strcmp(key, "gzip")) { h->gzip = value != 0; return SSC_OK; }
	if (!strcmp(key, "carry_pass2")) { h->carryPass2 = value != 0; return SSC_OK; }
	if (!strcmp(key, "concurrent_move")) {
		if (value < 0 || value > 64) return fail(SSC_ERR_INVALID, "concurrent_move: 0 (off) .. 64 SMs");
		h->moverCtas = (int)value;
		return SSC_OK;
	}
	if (!strcmp(key, "prefetch_windows")) { h->prefetchWindows = value != 0; return SSC_OK; }
	if (!strcmp(key, "no_q16")) {
		if (h->haveProfile) return fail(SSC_ERR_STATE, "no_q16 must be set before ssc_set_profile");
		h->noQ16 = value != 0;
		return SSC_OK;
	}
	if (!strcmp(key, "max_ctas")) { if (value < 0) return fail(SSC_ERR_INVALID, "max_ctas must be >= 0"); h->maxCtas = (int)value; return SSC_OK; }
	if (!strcmp(key, "no_splice")) { h->noSplice = value != 0; return SSC_OK; }
	if (!strcmp(key, "force_generic")) { h->forceGeneric = value != 0; h->slabPairs = 0; return SSC_OK; }
	if (!strcmp(key, "fp64_search")) {
		if (h->havePlan) return fail(SSC_ERR_STATE, "fp64_search must be set before ssc_set_plan");
		h->fp64 = value != 0;
		return SSC_OK;
	}
	return fail(SSC_ERR_INVALID, "unknown option %s", key);
}

int ssc_set_profile(ssc_handle* h, const ssc_profile_tables* t) {
	if (!h || !t) return fail(SSC_ERR_INVALID, "null argument");
	CK(cudaSetDevice(h->device));
	const char* err = ssc::build_tables(t, &h->th);
	if (err[0]) return fail(SSC_ERR_INVALID, "profile rejected: %s", err);
	if (t->read_length > 320) return fail(SSC_ERR_INVALID, "read_length %d > 320 is not supported", t->read_length);
	cudaStream_t s = h->compute;
	ssc::DeviceTablesHost& th = h->th;
	CK(h->d_isizeT.upload(th.isize.T, s)); CK(h->d_isizeSym.upload(th.isize.sym, s));
	CK(h->d_insT.upload(th.insLen.T, s)); CK(h->d_insSym.upload(th.insLen.sym, s));
	CK(h->d_delT.upload(th.delLen.T, s)); CK(h->d_delSym.upload(th.delLen.sym, s));
	CK(h->d_qualT.upload(th.qualT, s)); CK(h->d_qualSym.upload(th.qualSym, s));
	CK(h->d_qualDiagT.upload(th.qualDiagT, s)); CK(h->d_qualDiagSym.upload(th.qualDiagSym, s));
	std::vector<uint4> sub(th.sub.size());
	for (size_t i = 0; i < sub.size(); i++) sub[i] = make_uint4(th.sub[i].s0, th.sub[i].s1, th.sub[i].s2, th.sub[i].base);
	CK(h->d_sub.upload(sub, s));
	std::vector<int8_t> lut(th.asciiCode, th.asciiCode + 256);
	CK(h->d_lut.upload(lut, s));
	size_t nsub = (size_t)th.nRows * th.B * 4, nq = (size_t)16 * th.B * th.Q;
	auto up = [&](DevBuf<double>& d, const double* p, size_t n) -> cudaError_t {
		std::vector<double> v(p ? p : nullptr, p ? p + n : nullptr);
		return d.upload(v, s);
	};
	CK(up(h->d_fIsize, t->isize_cdf, t->n_isize > 0 ? t->n_isize : 0));
	CK(up(h->d_fIns, t->ins_cdf, t->n_ins));
	CK(up(h->d_fDel, t->del_cdf, t->n_del));
	CK(up(h->d_fSub1, t->subs_cdf1, nsub));
	CK(up(h->d_fSub2, t->use_cdf2 ? t->subs_cdf2 : nullptr, t->use_cdf2 ? nsub : 0));
	CK(up(h->d_fQual, t->quality_cdf, nq));
	CK(cudaStreamSynchronize(s));

	ssc::DevTables& d = h->dt;
	memset(&d, 0, sizeof(d));
	d.N = th.N; d.K = th.K; d.B = th.B; d.Q = th.Q; d.minQ = th.minQ; d.RL = th.RL; d.paired = th.paired;
	d.useCdf2 = th.useCdf2; d.fixedInsert = th.fixedInsert; d.minIS = th.minIS; d.nRows = th.nRows;
	d.insT = th.insT; d.delT = th.delT; d.insEnable = th.insEnable; d.delEnable = th.delEnable;
	d.nIsize = (int)th.isize.T.size(); d.nInsLen = (int)th.insLen.T.size(); d.nDelLen = (int)th.delLen.T.size();
	d.isizeT = h->d_isizeT.p; d.isizeSym = h->d_isizeSym.p;
	d.insLenT = h->d_insT.p; d.insLenSym = h->d_insSym.p;
	d.delLenT = h->d_delT.p; d.delLenSym = h->d_delSym.p;
	d.sub = h->d_sub.p; d.nSub = th.nRows * th.B;
	d.qualT = h->d_qualT.p; d.qualSym = h->d_qualSym.p; d.qualPitch = th.qualPitch; d.nQualRows = 16 * th.B;
	d.maxQualRow = th.maxQualRow; d.noQ16 = h->noQ16 ? 1 : 0;
	d.qualBins = ssc::fast_choose_qbins(th.B, th.RL);
	d.qualDiagT = h->d_qualDiagT.p; d.qualDiagSym = h->d_qualDiagSym.p; d.qualDiagPitch = th.diagPitch;
	d.compLut = th.compLut;
	d.baseChars = (uint32_t)(uint8_t)th.baseChar[0] | ((uint32_t)(uint8_t)th.baseChar[1] << 8) |
	              ((uint32_t)(uint8_t)th.baseChar[2] << 16) | ((uint32_t)(uint8_t)th.baseChar[3] << 24);
	d.f_isize = h->d_fIsize.p; d.f_ins = h->d_fIns.p; d.f_del = h->d_fDel.p;
	d.f_sub1 = h->d_fSub1.p; d.f_sub2 = h->d_fSub2.p; d.f_qual = h->d_fQual.p;
	d.f_nIsize = t->n_isize; d.f_nIns = t->n_ins; d.f_nDel = t->n_del;
	d.insertRate = t->insert_rate;
	volatile double one_minus = 1 - t->insert_rate;
	d.delThresh = t->del_rate / one_minus;
	{
		// the conflict-scored bin pitch of the shared quality image must not cost the all-rows-in-shared-memory mode
		int qm = 0; size_t sb = 0;
		if (d.qualPitch == 8 && !(ssc::fast_supported(d, h->smemLimit, &qm, &sb) && qm == 8)) d.qualBins = d.B;
	}
	h->haveProfile = true;
	h->havePlan = false;
	h->haveGz = false;
	return SSC_OK;
}

int ssc_genome_reserve(ssc_handle* h, uint64_t total_bases) {
	if (!h) return fail(SSC_ERR_INVALID, "null handle");
	if (!h->haveProfile) return fail(SSC_ERR_STATE, "ssc_set_profile must precede ssc_genome_reserve (codes follow the profile's base order)");
	CK(cudaSetDevice(h->device));
	uint64_t groups = (total_bases + SSC_GPAD + 31) / 32 + 2;
	CK(h->d_hap2.alloc(groups * 2 + 64));
	CK(h->d_hapN.alloc(groups + 64));
	CK(cudaMemsetAsync(h->d_hap2.p, 0, (groups * 2 + 64) * 4, h->compute));
	CK(cudaMemsetAsync(h->d_hapN.p, 0, (groups + 64) * 4, h->compute));
	for (int i = 0; i < 2; i++) {
		if (!h->h_stage[i]) CK(cudaMallocHost((void**)&h->h_stage[i], h->stageBytes));
		if (!h->d_stage[i]) CK(cudaMalloc((void**)&h->d_stage[i], h->stageBytes));
	}
	h->genomeCap = total_bases;
	h->genomeSize = 0;
	h->havePlan = false;
	return SSC_OK;
}

int ssc_genome_append(ssc_handle* h, const char* ascii, uint64_t n, uint64_t* first_base) {
	if (!h || (!ascii && n)) return fail(SSC_ERR_INVALID, "null argument");
	if (h->genomeSize + n > h->genomeCap) return fail(SSC_ERR_INVALID, "genome append exceeds the reserved %llu bases", (unsigned long long)h->genomeCap);
	CK(cudaSetDevice(h->device));
	if (first_base) *first_base = h->genomeSize;
	uint64_t done = 0;
	int k = 0;
	while (done < n) {
		uint64_t chunk = std::min<uint64_t>(h->stageBytes, n - done);
		CK(cudaEventSynchronize(h->evStage[k]));           // staging buffer k free again
		memcpy(h->h_stage[k], ascii + done, chunk);
		CK(cudaMemcpyAsync(h->d_stage[k], h->h_stage[k], chunk, cudaMemcpyHostToDevice, h->compute));
		CK(ssc::launch_pack(h->d_stage[k], chunk, SSC_GPAD + h->genomeSize + done, h->d_hap2.p, h->d_hapN.p, h->d_lut.p, h->compute));
		CK(cudaEventRecord(h->evStage[k], h->compute));
		h->stats.launches += 1;
		h->stats.h2d_bytes += chunk;
		done += chunk;
		k ^= 1;
	}
	h->genomeSize += n;
	return SSC_OK;
}

int ssc_reference_upload(ssc_handle* h, const char* ascii, uint64_t n) {
	if (!h || (!ascii && n)) return fail(SSC_ERR_INVALID, "null argument");
	CK(cudaSetDevice(h->device));
	if (h->d_ref.n < n) CK(h->d_ref.alloc((size_t)n + (size_t)n / 8 + 4096));
	CK(cudaStreamSynchronize(h->compute));          // earlier ssc_genome_append_ref launches still read the old reference
	// in stream order with the pack kernels that read it: a plain cudaMemcpy from pageable memory runs on the legacy stream,
	// which the (non-blocking) compute stream does not wait for, and may return before the DMA has landed
	if (n) CK(cudaMemcpyAsync(h->d_ref.p, ascii, (size_t)n, cudaMemcpyHostToDevice, h->compute));
	CK(cudaStreamSynchronize(h->compute));          // the caller may free the chromosome string
	h->refSize = n;
	h->stats.h2d_bytes += n;
	return SSC_OK;
}

}  // extern "C"

namespace {
// One FASTA record: file -> pinned staging (pread) -> raw (device) -> unfold kernel -> ref (device); returns when the
// unfolded record and the count of IUPAC characters are there.  `slot` 0 = the foreground set of staging buffers, 1 = the
// prefetch set.  Thread-safe against a concurrent call on the other slot (nothing of the handle is shared but the device).
int upload_fasta_record(ssc_handle* h, int slot, cudaStream_t s, DevBuf<uint8_t>& ref, DevBuf<uint8_t>& raw, DevBuf<unsigned long long>& otherBuf,
                        cudaEvent_t waitFor, int fd, uint64_t file_offset, uint64_t raw_len, uint64_t n_bases, uint32_t line_bases,
                        uint32_t line_width, uint64_t* n_other) {
	CK(cudaSetDevice(h->device));
	if (ref.n < n_bases + 16) CK(ref.alloc((size_t)n_bases + (size_t)n_bases / 8 + 4096));
	if (raw.n < raw_len) CK(raw.alloc((size_t)raw_len + (size_t)raw_len / 8 + 4096));
	if (!otherBuf.p) CK(otherBuf.alloc(1));
	const int NR = ssc_handle::FA_READERS, b0 = slot * NR;
	for (int i = b0; i < b0 + NR; i++) {
		if (!h->h_fa[i]) CK(cudaMallocHost((void**)&h->h_fa[i], h->faBytes));
		if (!h->evFa[i]) CK(cudaEventCreateWithFlags(&h->evFa[i], cudaEventDisableTiming));
	}
	if (waitFor) CK(cudaStreamWaitEvent(s, waitFor, 0));         // earlier readers of `ref` (pack kernels of the record before the last)
	CK(cudaMemsetAsync(otherBuf.p, 0, 8, s));
	// NR staging buffers, each fed by its own host thread (chunk c goes to reader c mod NR): the page-cache copy of pread is the
	// slow part, so two run at a time, under the DMA of earlier chunks
	const uint64_t nChunks = (raw_len + h->faBytes - 1) / h->faBytes;
	int rcs[ssc_handle::FA_READERS] = {SSC_OK, SSC_OK};
	std::string errMsg[ssc_handle::FA_READERS];
	auto feed = [&](int k) {
		if (cudaSetDevice(h->device) != cudaSuccess) { rcs[k] = SSC_ERR_CUDA; errMsg[k] = "cudaSetDevice failed"; return; }
		uint8_t* buf = h->h_fa[b0 + k];
		cudaEvent_t ev = h->evFa[b0 + k];
		for (uint64_t c = (uint64_t)k; c < nChunks; c += NR) {
			const uint64_t done = c * h->faBytes;
			const size_t chunk = (size_t)std::min<uint64_t>(h->faBytes, raw_len - done);
			cudaError_t e = cudaEventSynchronize(ev);                     // staging buffer free again
			size_t got = 0;
			while (e == cudaSuccess && got < chunk) {
				const ssize_t r = pread(fd, buf + got, chunk - got, (off_t)(file_offset + done + got));
				if (r < 0) { if (errno == EINTR) continue; rcs[k] = SSC_ERR_INVALID; errMsg[k] = std::string("reading the FASTA file failed: ") + strerror(errno); return; }
				if (r == 0) break;                                         // a last line without a line feed ends the file early
				got += (size_t)r;
			}
			if (got < chunk) memset(buf + got, '\n', chunk - got);
			if (e == cudaSuccess) e = cudaMemcpyAsync(raw.p + done, buf, chunk, cudaMemcpyHostToDevice, s);
			if (e == cudaSuccess) e = cudaEventRecord(ev, s);
			if (e != cudaSuccess) { rcs[k] = SSC_ERR_CUDA; errMsg[k] = std::string("staging the FASTA record failed: ") + cudaGetErrorString(e); return; }
		}
	};
	{
		std::vector<std::thread> readers;
		for (int k = 1; k < NR && (uint64_t)k < nChunks; k++) readers.emplace_back(feed, k);
		feed(0);
		for (auto& t : readers) t.join();
	}
	for (int k = 0; k < NR; k++) if (rcs[k]) return fail(rcs[k], "%s", errMsg[k].c_str());
	CK(ssc::launch_unfold(raw.p, raw_len, n_bases, line_bases, line_width, ref.p, otherBuf.p, s));
	unsigned long long other = 0;
	CK(cudaMemcpyAsync(&other, otherBuf.p, 8, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	if (n_other) *n_other = other;
	return SSC_OK;
}

int check_fasta_geometry(uint64_t n_bases, uint32_t line_bases, uint32_t line_width) {
	if (n_bases >= (1ull << 32) - 64 || (n_bases && (line_bases == 0 || line_width < line_bases)))
		return fail(SSC_ERR_INVALID, "FASTA record geometry not supported (%llu bases, %u per line of %u bytes)",
		            (unsigned long long)n_bases, line_bases, line_width);
	return SSC_OK;
}
}  // namespace

extern "C" {

int ssc_reference_upload_fasta(ssc_handle* h, int fd, uint64_t file_offset, uint64_t raw_len, uint64_t n_bases,
                               uint32_t line_bases, uint32_t line_width, uint64_t* n_other) {
	if (!h || fd < 0) return fail(SSC_ERR_INVALID, "bad argument");
	if (h->prefetching) return fail(SSC_ERR_STATE, "a prefetched FASTA record is waiting: adopt it first");
	int rc = check_fasta_geometry(n_bases, line_bases, line_width);
	if (rc) return rc;
	rc = upload_fasta_record(h, 0, h->compute, h->d_ref, h->d_raw, h->d_other, nullptr, fd, file_offset, raw_len, n_bases, line_bases, line_width, n_other);
	if (rc) return rc;
	h->refSize = n_bases;
	h->stats.h2d_bytes += raw_len;
	h->stats.launches += 1;
	return SSC_OK;
}

int ssc_reference_prefetch_fasta(ssc_handle* h, int fd, uint64_t file_offset, uint64_t raw_len, uint64_t n_bases,
                                 uint32_t line_bases, uint32_t line_width) {
	if (!h || fd < 0) return fail(SSC_ERR_INVALID, "bad argument");
	if (h->prefetching) return fail(SSC_ERR_STATE, "one FASTA record can be prefetched at a time");
	int rc = check_fasta_geometry(n_bases, line_bases, line_width);
	if (rc) return rc;
	CK(cudaSetDevice(h->device));
	if (!h->upload) CK(cudaStreamCreateWithFlags(&h->upload, cudaStreamNonBlocking));
	if (!h->evRefFree) CK(cudaEventCreateWithFlags(&h->evRefFree, cudaEventDisableTiming));
	// the second buffer set was the current reference until the last adoption: its readers are in the compute stream by now
	CK(cudaEventRecord(h->evRefFree, h->compute));
	h->prefetching = true;
	h->prefetchRc = 0; h->prefetchErr.clear();
	h->prefetchBases = n_bases; h->prefetchRaw = raw_len; h->prefetchOther = 0;
	h->prefetchThread = std::thread([=]() {
		h->prefetchRc = upload_fasta_record(h, 1, h->upload, h->d_ref2, h->d_raw2, h->d_other2, h->evRefFree, fd, file_offset, raw_len, n_bases,
		                                    line_bases, line_width, &h->prefetchOther);
		if (h->prefetchRc) h->prefetchErr = g_err;          // (this thread's message, for the adopting thread)
	});
	return SSC_OK;
}

int ssc_reference_adopt_prefetched(ssc_handle* h, uint64_t* n_other) {
	if (!h) return fail(SSC_ERR_INVALID, "null handle");
	if (!h->prefetching) return fail(SSC_ERR_STATE, "no FASTA record was prefetched");
	h->prefetchThread.join();
	h->prefetching = false;
	if (h->prefetchRc) return fail(h->prefetchRc, "%s", h->prefetchErr.c_str());
	std::swap(h->d_ref, h->d_ref2); std::swap(h->d_raw, h->d_raw2); std::swap(h->d_other, h->d_other2);
	h->refSize = h->prefetchBases;
	h->stats.h2d_bytes += h->prefetchRaw;
	h->stats.launches += 1;
	if (n_other) *n_other = h->prefetchOther;
	return SSC_OK;
}

int ssc_genome_append_ref(ssc_handle* h, uint64_t ref_off, uint64_t len, int32_t reps, uint64_t* first_base) {
	if (!h || reps < 0) return fail(SSC_ERR_INVALID, "bad argument");
	if (ref_off + len > h->refSize) return fail(SSC_ERR_INVALID, "reference range outside the uploaded chromosome");
	if (h->genomeSize + len * (uint64_t)reps > h->genomeCap) return fail(SSC_ERR_INVALID, "genome append exceeds the reserved %llu bases", (unsigned long long)h->genomeCap);
	CK(cudaSetDevice(h->device));
	if (first_base) *first_base = h->genomeSize;
	for (int32_t r = 0; r < reps && len; r++) {
		CK(ssc::launch_pack(h->d_ref.p + ref_off, len, SSC_GPAD + h->genomeSize, h->d_hap2.p, h->d_hapN.p, h->d_lut.p, h->compute));
		h->genomeSize += len;
		h->stats.launches += 1;
	}
	return SSC_OK;
}

int ssc_genome_poke(ssc_handle* h, const int64_t* store_pos, const char* chars, int64_t n) {
	if (!h || n < 0 || (n && (!store_pos || !chars))) return fail(SSC_ERR_INVALID, "null argument");
	if (n == 0) return SSC_OK;
	CK(cudaSetDevice(h->device));
	for (int64_t i = 0; i < n; i++)
		if (store_pos[i] < 0 || (uint64_t)store_pos[i] >= h->genomeSize) return fail(SSC_ERR_INVALID, "poke %lld outside the haplotype store", (long long)i);
	DevBuf<int64_t> d_pos; DevBuf<uint8_t> d_ch;
	CK(d_pos.alloc((size_t)n)); CK(d_ch.alloc((size_t)n));
	CK(cudaMemcpyAsync(d_pos.p, store_pos, (size_t)n * 8, cudaMemcpyHostToDevice, h->compute));
	CK(cudaMemcpyAsync(d_ch.p, chars, (size_t)n, cudaMemcpyHostToDevice, h->compute));
	CK(ssc::launch_poke(h->d_hap2.p, h->d_hapN.p, d_pos.p, d_ch.p, n, h->d_lut.p, h->compute));
	CK(cudaStreamSynchronize(h->compute));
	d_pos.release(); d_ch.release();
	h->stats.launches += 1;
	return SSC_OK;
}

int ssc_genome_read(ssc_handle* h, uint64_t start, uint64_t n, char* out) {
	if (!h || (!out && n)) return fail(SSC_ERR_INVALID, "null argument");
	if (start + n > h->genomeSize) return fail(SSC_ERR_INVALID, "range outside the haplotype store");
	if (n == 0) return SSC_OK;
	CK(cudaSetDevice(h->device));
	DevBuf<uint8_t> d_out;
	CK(d_out.alloc((size_t)n));
	CK(ssc::launch_unpack(h->d_hap2.p, h->d_hapN.p, SSC_GPAD + start, n, h->dt.baseChars, d_out.p, h->compute));
	CK(cudaMemcpyAsync(out, d_out.p, (size_t)n, cudaMemcpyDeviceToHost, h->compute));
	CK(cudaStreamSynchronize(h->compute));
	d_out.release();
	h->stats.launches += 1;
	return SSC_OK;
}

int ssc_genome_size(ssc_handle* h, uint64_t* n_bases) {
	if (!h || !n_bases) return fail(SSC_ERR_INVALID, "null argument");
	*n_bases = h->genomeSize;
	return SSC_OK;
}

int ssc_gc_census(ssc_handle* h, const int64_t* starts, const int32_t* lens, int64_t n, int32_t* gc, int32_t* nn) {
	if (!h || n < 0 || (n && (!starts || !lens || !gc || !nn))) return fail(SSC_ERR_INVALID, "null argument");
	if (!h->haveProfile || !h->d_hap2.p) return fail(SSC_ERR_STATE, "ssc_genome_append must precede ssc_gc_census");
	if (n == 0) return SSC_OK;
	if (n > 0x7fffffffLL * 8) return fail(SSC_ERR_INVALID, "too many intervals");
	CK(cudaSetDevice(h->device));
	for (int64_t i = 0; i < n; i++)
		if (starts[i] < 0 || lens[i] < 0 || (uint64_t)(starts[i] + lens[i]) > h->genomeSize)
			return fail(SSC_ERR_INVALID, "interval %lld outside the haplotype store", (long long)i);
	uint32_t gcCodes = 0;
	for (int k = 0; k < 4; k++) if (h->th.baseChar[k] == 'G' || h->th.baseChar[k] == 'C') gcCodes |= 1u << k;
	cudaStream_t s = h->compute;
	const bool timing = getenv("SIMUSCOP_TIMING") != nullptr;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t0 = now();
	if (timing) { cudaStreamSynchronize(s); fprintf(stderr, "[simuscop timing] gc_census: drained the upload queue in %.3f s\n", now() - t0); t0 = now(); }
	// census buffers live in the handle and only grow (a cudaMalloc/cudaFree pair per call stalls for up to 0.7 s)
	DevBuf<int64_t>& d_starts = h->d_cenStarts; DevBuf<int32_t>& d_lens = h->d_cenLens; DevBuf<int32_t>& d_gc = h->d_cenGc; DevBuf<int32_t>& d_nn = h->d_cenNn;
	if (d_starts.n < (size_t)n) {
		const size_t cap = (size_t)n + (size_t)n / 2 + 1024;
		CK(d_starts.alloc(cap)); CK(d_lens.alloc(cap)); CK(d_gc.alloc(cap)); CK(d_nn.alloc(cap));
	}
	CK(cudaMemcpyAsync(d_starts.p, starts, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(d_lens.p, lens, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	CK(ssc::launch_gc_census(h->d_hap2.p, h->d_hapN.p, d_starts.p, d_lens.p, n, gcCodes, d_gc.p, d_nn.p, s));
	CK(cudaMemcpyAsync(gc, d_gc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(nn, d_nn.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	if (timing) fprintf(stderr, "[simuscop timing] gc_census: %lld intervals in %.3f s\n", (long long)n, now() - t0);
	h->stats.launches += 1;
	h->stats.h2d_bytes += (uint64_t)n * 12; h->stats.d2h_bytes += (uint64_t)n * 8;
	return SSC_OK;
}

int ssc_set_plan(ssc_handle* h, uint64_t seed, const ssc_bin* bins, int64_t n_bins, const ssc_segment* segs,
                 int64_t n_segs, const char* names, int64_t names_len, int64_t* planned_pairs, int64_t* emitted_pairs) {
	if (!h || (!bins && n_bins) || (!segs && n_segs)) return fail(SSC_ERR_INVALID, "null argument");
	if (!h->haveProfile) return fail(SSC_ERR_STATE, "ssc_set_profile must precede ssc_set_plan");
	CK(cudaSetDevice(h->device));
	const ssc::DevTables& t = h->dt;
	const int RL = t.RL;
	const bool paired = t.paired != 0;
	h->seed = seed;
	h->havePlan = false;
	h->haveGz = false;
	h->pending.valid = false;

	// The per-bin passes below run on a few host threads over contiguous bin / segment ranges (a 3 Gb diploid plan has
	// 6 M bins); every result is position-determined, so the plan does not depend on the thread count.
	const unsigned T = (unsigned)std::max<int64_t>(1, std::min<int64_t>({16, (int64_t)std::thread::hardware_concurrency(), n_bins / 65536 + 1}));
	auto parallel = [&](int64_t n, const std::function<void(int64_t, int64_t, unsigned)>& fn) {
		if (T == 1 || n < 2) { fn(0, n, 0); return; }
		std::vector<std::thread> th;
		for (unsigned k = 1; k < T; k++) th.emplace_back(fn, n * k / T, n * (k + 1) / T, k);
		fn(0, n / T, 0);
		for (auto& x : th) x.join();
	};

	// ---- validate, planned pairs, risky bins
	std::vector<int64_t>& pb = h->planBaseAll;
	pb.assign(n_bins + 1, 0);
	std::vector<int32_t> planned(n_bins), emit(n_bins);
	std::vector<int32_t> riskyBase(n_bins, -1);
	int maxName = 0;
	if (names_len < 0 || names_len >= (1 << 17)) return fail(SSC_ERR_INVALID, "name blob of %lld bytes (limit 131071)", (long long)names_len);
	for (int64_t s = 0; s < n_segs; s++) {
		if (segs[s].first_bin < 0 || segs[s].n_bins < 0 || segs[s].first_bin + segs[s].n_bins > n_bins)
			return fail(SSC_ERR_INVALID, "segment %lld: bin range out of bounds", (long long)s);
		if (segs[s].name_len < 0 || segs[s].name_len > 64 || segs[s].name_offset < 0 || segs[s].name_offset + segs[s].name_len > names_len)
			return fail(SSC_ERR_INVALID, "segment %lld: bad name (max 64 bytes)", (long long)s);
		maxName = std::max(maxName, (int)segs[s].name_len);
	}
	struct BinError { int64_t bin = -1; const char* what = nullptr; };
	std::vector<BinError> errs(T);
	std::vector<std::vector<int64_t>> riskyOf(T);              // bins that can fail, per thread range (ascending)
	parallel(n_bins, [&](int64_t lo, int64_t hi, unsigned tid) {
		for (int64_t i = lo; i < hi; i++) {
			const ssc_bin& b = bins[i];
			const int64_t n = b.read_count > 0 ? (paired ? ((int64_t)b.read_count + 1) / 2 : (int64_t)b.read_count) : 0;
			if (n > 0 && errs[tid].bin < 0) {
				const char* what = nullptr;
				if (b.segment < 0 || b.segment >= n_segs) what = "bad segment";
				else if (b.spos < 0 || b.epos < b.spos) what = "bad range";
				else if (b.segsize == 0) what = "segsize 0 (copy number 0 segments cannot emit reads)";
				else if (b.hap_base < 0 || b.contig_end < b.hap_base || (uint64_t)b.contig_end > h->genomeSize) what = "haplotype range outside the store";
				else if (n > 0x7fffffff) what = "too many reads";
				if (what) { errs[tid].bin = i; errs[tid].what = what; }
			}
			planned[i] = (int32_t)n;
			emit[i] = (int32_t)n;
			if (n > 0) {
				// an attempt fails iff min(want, contig_end - (hap_base+pos)) < RL  (Segment.cpp:753)
				bool risky = (int64_t)b.epos > b.contig_end - b.hap_base - RL;
				if (!paired && ((int64_t)b.epos - b.spos + 1) < RL) risky = true;
				if (paired && t.nIsize == 0 && t.fixedInsert < RL) risky = true;
				if (paired && t.nIsize > 0 && t.minIS < RL) risky = true;          // an insert-size table that reaches below the read length
				if (risky) riskyOf[tid].push_back(i);
			}
		}
	});
	for (unsigned k = 0; k < T; k++)
		if (errs[k].bin >= 0) return fail(SSC_ERR_INVALID, "bin %lld: %s", (long long)errs[k].bin, errs[k].what);
	for (int64_t i = 0; i < n_bins; i++) pb[i + 1] = pb[i] + planned[i];
	h->plannedPairs = pb[n_bins];
	std::vector<ssc::CensusBin> census;
	std::vector<int64_t> censusOf;            // census index -> bin
	int64_t riskyTotal = 0;
	for (unsigned k = 0; k < T; k++)
		for (int64_t i : riskyOf[k]) {
			const ssc_bin& b = bins[i];
			const int64_t n = planned[i];
			ssc::CensusBin c;
			c.hap_base = b.hap_base + SSC_GPAD; c.contig_end = b.contig_end + SSC_GPAD; c.plan_base = pb[i];
			c.spos = b.spos; c.epos = b.epos; c.planned = (int32_t)n; c.risky_base = (int32_t)riskyTotal;
			if (riskyTotal + n > 0x7fffffff) return fail(SSC_ERR_INVALID, "too many pairs in bins that can fail");
			riskyBase[i] = (int32_t)riskyTotal;
			riskyTotal += n;
			census.push_back(c);
			censusOf.push_back(i);
		}

	// ---- census on the device
	cudaStream_t s = h->compute;
	CK(h->d_risky.alloc((size_t)std::max<int64_t>(riskyTotal, 1)));
	if (!census.empty()) {
		DevBuf<ssc::CensusBin> d_census;
		DevBuf<int32_t> d_emitted;
		CK(d_census.upload(census, s));
		CK(d_emitted.alloc(census.size()));
		CK(ssc::launch_census(t, h->fp64, d_census.p, (int)census.size(), seed, h->d_risky.p, d_emitted.p, s));
		std::vector<int32_t> em(census.size());
		CK(cudaMemcpyAsync(em.data(), d_emitted.p, em.size() * 4, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		h->stats.launches += 1;
		for (size_t k = 0; k < census.size(); k++) emit[censusOf[k]] = em[k];
		d_census.release(); d_emitted.release();
	}

	// ---- prefix sums, compaction
	std::vector<int64_t>& eb = h->emitBaseAll;
	eb.assign(n_bins + 1, 0);
	for (int64_t i = 0; i < n_bins; i++) eb[i + 1] = eb[i] + emit[i];
	h->emittedPairs = eb[n_bins];
	// device bins of segment sIdx go to dev[segOff[sIdx] ...): bins must be listed segment by segment in order, otherwise
	// emit order != bin order
	std::vector<int64_t> segOff(n_segs + 1, 0);
	std::vector<const char*> segErr(T, nullptr);
	std::vector<int64_t> segErrAt(T, -1);
	parallel(n_segs, [&](int64_t lo, int64_t hi, unsigned tid) {
		for (int64_t sIdx = lo; sIdx < hi; sIdx++) {
			int64_t cnt = 0, fragBase = 0;
			for (int64_t i = segs[sIdx].first_bin; i < segs[sIdx].first_bin + segs[sIdx].n_bins; i++) {
				if (emit[i] > 0) {
					cnt++;
					if (segErrAt[tid] < 0 && bins[i].segment != sIdx) { segErrAt[tid] = i; segErr[tid] = "does not belong to the segment that lists it"; }
					if (segErrAt[tid] < 0 && fragBase + emit[i] > 0x7fffffff) { segErrAt[tid] = i; segErr[tid] = "fragment counter overflow of its segment"; }
				}
				fragBase += emit[i];
			}
			segOff[sIdx + 1] = cnt;
		}
	});
	for (unsigned k = 0; k < T; k++) if (segErrAt[k] >= 0) return fail(SSC_ERR_INVALID, "bin %lld %s", (long long)segErrAt[k], segErr[k]);
	for (int64_t sIdx = 0; sIdx < n_segs; sIdx++) segOff[sIdx + 1] += segOff[sIdx];
	const size_t nDev = (size_t)segOff[n_segs];
	std::vector<int64_t>& devEmitBase = h->devEmitBase;
	devEmitBase.assign(nDev + 1, 0);
	// The device bins (64 bytes each, 384 MB for a 3 Gb diploid plan) are formatted straight into the two pinned staging
	// buffers, a run of whole segments at a time, and copied from there: no pageable image of the table is ever built.
	for (int i = 0; i < 2; i++) {
		if (!h->h_stage[i]) CK(cudaMallocHost((void**)&h->h_stage[i], h->stageBytes));
		CK(cudaEventSynchronize(h->evStage[i]));
	}
	CK(h->d_bins.alloc(std::max<size_t>(nDev, 1)));
	const size_t perChunk = h->stageBytes / sizeof(ssc::DevBin);
	int stage = 0;
	for (int64_t s0 = 0; s0 < n_segs;) {
		int64_t s1 = s0;
		while (s1 < n_segs && (size_t)(segOff[s1 + 1] - segOff[s0]) <= perChunk) s1++;
		if (s1 == s0) return fail(SSC_ERR_INVALID, "segment %lld has more than %zu bins", (long long)s0, perChunk);
		const int64_t o0 = segOff[s0], cnt = segOff[s1] - o0;
		if (cnt > 0) {
			CK(cudaEventSynchronize(h->evStage[stage]));
			ssc::DevBin* out = (ssc::DevBin*)h->h_stage[stage];
			parallel(s1 - s0, [&](int64_t lo, int64_t hi, unsigned) {
				for (int64_t sIdx = s0 + lo; sIdx < s0 + hi; sIdx++) {
					int64_t fragBase = 0, o = segOff[sIdx];
					for (int64_t i = segs[sIdx].first_bin; i < segs[sIdx].first_bin + segs[sIdx].n_bins; i++) {
						if (emit[i] > 0) {
							ssc::DevBin d;
							d.hap_base = bins[i].hap_base + SSC_GPAD; d.contig_end = bins[i].contig_end + SSC_GPAD;
							d.plan_base = pb[i]; d.emit_base = eb[i];
							d.spos = bins[i].spos; d.epos = bins[i].epos; d.segsize = bins[i].segsize;
							d.frag_base = (int32_t)fragBase;
							d.name_off = segs[sIdx].name_offset; d.name_len = segs[sIdx].name_len;
							d.risky_base = riskyBase[i]; d.pad = 0;
							out[o - o0] = d;
							devEmitBase[(size_t)o] = eb[i];
							o++;
						}
						fragBase += emit[i];
					}
				}
			});
			CK(cudaMemcpyAsync(h->d_bins.p + o0, out, (size_t)cnt * sizeof(ssc::DevBin), cudaMemcpyHostToDevice, s));
			CK(cudaEventRecord(h->evStage[stage], s));
			stage ^= 1;
		}
		s0 = s1;
	}
	devEmitBase[nDev] = h->emittedPairs;
	for (size_t k = 1; k + 1 < devEmitBase.size(); k++)
		if (devEmitBase[k] <= devEmitBase[k - 1]) return fail(SSC_ERR_INVALID, "segments must list their bins in increasing, non-overlapping order");
	h->nDevBins = (int64_t)nDev;
	CK(h->d_emitBase.upload(devEmitBase, s));
	std::vector<char> nm(names, names + names_len);
	nm.push_back(0);
	CK(h->d_names.upload(nm, s));
	CK(cudaStreamSynchronize(s));
	h->stats.h2d_bytes += nDev * sizeof(ssc::DevBin) + devEmitBase.size() * 8;
	// worst-case record bytes per pair per file: header + 2*(RL + 96) + 4
	h->maxRecBytes = maxName + 10 + 1 + 10 + 3 + 2 * (RL + 16) + 4;
	h->havePlan = true;
	if (planned_pairs) *planned_pairs = h->plannedPairs;
	if (emitted_pairs) *emitted_pairs = h->emittedPairs;
	return SSC_OK;
}

int ssc_generate(ssc_handle* h, int64_t pair_lo, int64_t pair_hi, ssc_sink_fn sink, void* user) {
	if (!h || !sink) return fail(SSC_ERR_INVALID, "null argument");
	if (!h->havePlan) return fail(SSC_ERR_STATE, "ssc_set_plan must precede ssc_generate");
	if (pair_lo < 0 || pair_hi < pair_lo) return fail(SSC_ERR_INVALID, "bad pair range");
	CK(cudaSetDevice(h->device));
	int rc = ensure_batch_resources(h, true);
	if (rc) return rc;
	const int64_t eLo = emit_index_of_plan(h, pair_lo), eHi = emit_index_of_plan(h, pair_hi);
	if (eHi <= eLo) return SSC_OK;
	const int nFiles = h->dt.paired ? 2 : 1;
	struct Batch { int64_t lo, hi; };
	std::vector<Batch> batches;
	for (int64_t e = eLo; e < eHi; e += h->slabPairs) batches.push_back({e, std::min(eHi, e + h->slabPairs)});
	const int nb = (int)batches.size();
	// pipeline: the launch of batch k+1 (which also completes the dense slab of batch k) runs under the device->host copies of
	// batch k-1 (file 1 / file 2 on two streams), which run under the sink of batch k-2.  h_out[k & 1] was consumed by the
	// sink of batch k-2 an iteration before the copy of batch k is issued.
	int qm = 0; size_t sb = 0;
	const bool carried = use_fast(h, &qm, &sb) && !h->gzip && (h->carryPass2 || h->moverCtas > 0);     // pass 2b of batch k rides on the launch of batch k+1
	auto launch = [&](int k) -> int {
		// the dense slab this launch writes: d_out[(k-1) & 1] when it carries the previous batch's moves, else d_out[k & 1];
		// its last reader is the copy of the batch two before the one that now lands there
		const int target = carried ? (k + 1) & 1 : k & 1;
		CK(cudaStreamWaitEvent(h->compute, h->evCopy[target], 0));
		CK(cudaStreamWaitEvent(h->compute, h->evCopy2[target], 0));
		int r = launch_batch(h, k & 1, batches[k].lo, batches[k].hi);
		if (r) return r;
		CK(cudaEventRecord(h->evGen[k & 1], h->compute));
		if (!carried) CK(cudaEventRecord(h->evDense[k & 1], h->compute));     // the dense slab of batch k is complete with its own launch
		return SSC_OK;
	};
	rc = launch(0);
	if (rc) return bail(h, rc);
	ssc::BatchResult res[2];
	for (int k = 0; k < nb; k++) {
		const int buf = k & 1;
		rc = k + 1 < nb ? launch(k + 1) : flush_pending(h);         // after this the dense slab of batch k is in the stream
		if (rc) return bail(h, rc);
		if (carried) CK(cudaEventRecord(h->evDense[buf], h->compute));    // (its pass 2b came with the launch of batch k+1)
		CK(cudaEventSynchronize(h->evGen[buf]));
		res[buf] = *h->h_result[buf];
		rc = check_result(h, res[buf]);
		if (rc) return bail(h, rc);
		CK(cudaStreamWaitEvent(h->copy, h->evDense[buf], 0));
		CK(cudaMemcpyAsync(h->h_out[buf][0], h->d_out[buf][0], res[buf].bytes1, cudaMemcpyDeviceToHost, h->copy));
		CK(cudaEventRecord(h->evCopy[buf], h->copy));
		if (nFiles == 2) {
			CK(cudaStreamWaitEvent(h->copy2, h->evDense[buf], 0));
			CK(cudaMemcpyAsync(h->h_out[buf][1], h->d_out[buf][1], res[buf].bytes2, cudaMemcpyDeviceToHost, h->copy2));
		}
		CK(cudaEventRecord(h->evCopy2[buf], h->copy2));
		h->stats.d2h_bytes += res[buf].bytes1 + res[buf].bytes2;
		if (k >= 1) {
			const int pb = (k - 1) & 1;
			CK(cudaEventSynchronize(h->evCopy[pb]));
			CK(cudaEventSynchronize(h->evCopy2[pb]));
			int src = sink(user, (const char*)h->h_out[pb][0], res[pb].bytes1, nFiles == 2 ? (const char*)h->h_out[pb][1] : nullptr,
			               nFiles == 2 ? res[pb].bytes2 : 0, batches[k - 1].lo, batches[k - 1].hi - batches[k - 1].lo);
			if (src) { fail(SSC_ERR_SINK, "sink returned %d", src); return bail(h, SSC_ERR_SINK); }
		}
	}
	{
		const int pb = (nb - 1) & 1;
		CK(cudaEventSynchronize(h->evCopy[pb]));
		CK(cudaEventSynchronize(h->evCopy2[pb]));
		int src = sink(user, (const char*)h->h_out[pb][0], res[pb].bytes1, nFiles == 2 ? (const char*)h->h_out[pb][1] : nullptr,
		               nFiles == 2 ? res[pb].bytes2 : 0, batches[nb - 1].lo, batches[nb - 1].hi - batches[nb - 1].lo);
		if (src) return fail(SSC_ERR_SINK, "sink returned %d", src);
	}
	return SSC_OK;
}

int ssc_generate_device(ssc_handle* h, int64_t pair_lo, int64_t pair_hi, uint64_t* bytes1, uint64_t* bytes2,
                        uint64_t* bases, double* device_ms) {
	if (!h) return fail(SSC_ERR_INVALID, "null handle");
	if (!h->havePlan) return fail(SSC_ERR_STATE, "ssc_set_plan must precede ssc_generate_device");
	if (pair_lo < 0 || pair_hi < pair_lo) return fail(SSC_ERR_INVALID, "bad pair range");
	CK(cudaSetDevice(h->device));
	int rc = ensure_batch_resources(h, false);
	if (rc) return rc;
	const int64_t eLo = emit_index_of_plan(h, pair_lo), eHi = emit_index_of_plan(h, pair_hi);
	uint64_t b1 = 0, b2 = 0, nb = 0;
	float ms = 0;
	if (eHi > eLo) {
		h->timeKernels = true; h->kevUsed = 0;
		CK(cudaEventRecord(h->evStart, h->compute));
		int k = 0;
		for (int64_t e = eLo; e < eHi; e += h->slabPairs, k++) {
			const int buf = k & 1;
			if (k >= 2) {
				CK(cudaEventSynchronize(h->evGen[buf]));
				ssc::BatchResult r = *h->h_result[buf];
				rc = check_result(h, r);
				if (rc) return bail(h, rc);
				b1 += r.bytes1; b2 += r.bytes2; nb += r.bases;
			}
			rc = launch_batch(h, buf, e, std::min(eHi, e + h->slabPairs));
			if (rc) return bail(h, rc);
			CK(cudaEventRecord(h->evGen[buf], h->compute));
		}
		// the last batch's blobs: stand-alone move, timed with the scans as "pass 2"
		while ((int)h->kev.size() < h->kevUsed + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->kev.push_back(e); }
		const int fl = h->kevUsed;
		CK(cudaEventRecord(h->kev[fl], h->compute));
		rc = flush_pending(h);
		if (rc) return bail(h, rc);
		CK(cudaEventRecord(h->kev[fl + 1], h->compute));
		CK(cudaEventRecord(h->evStop, h->compute));
		CK(cudaEventSynchronize(h->evStop));
		for (int j = std::max(0, k - 2); j < k; j++) {
			ssc::BatchResult r = *h->h_result[j & 1];
			rc = check_result(h, r);
			if (rc) return rc;
			b1 += r.bytes1; b2 += r.bytes2; nb += r.bases;
		}
		CK(cudaEventElapsedTime(&ms, h->evStart, h->evStop));
		h->stats.device_ms += ms;
		h->timeKernels = false;
		for (int i = 0; i + 2 < h->kevUsed; i += 3) {
			float a = 0, b = 0;
			CK(cudaEventElapsedTime(&a, h->kev[i], h->kev[i + 1]));
			CK(cudaEventElapsedTime(&b, h->kev[i + 1], h->kev[i + 2]));
			h->stats.gen_kernel_ms += a; h->stats.compact_kernel_ms += b; h->stats.timed_batches += 1;
		}
		{
			float c = 0;
			CK(cudaEventElapsedTime(&c, h->kev[fl], h->kev[fl + 1]));
			h->stats.compact_kernel_ms += c;
		}
	}
	if (bytes1) *bytes1 = b1;
	if (bytes2) *bytes2 = b2;
	if (bases) *bases = nb;
	if (device_ms) *device_ms = ms;
	return SSC_OK;
}

int ssc_table_lookup_host(const double* cdf, int n, uint32_t u) {
	if (!cdf || n < 1) return -1;
	ssc::CompressedCdf c = ssc::compress_cdf(cdf, n);
	size_t k = 0;
	while (k + 1 < c.T.size() && c.T[k] < u) k++;
	return (int)c.sym[k];
}

int ssc_sub_lookup_host(const double* cdf4, uint32_t u) {
	if (!cdf4) return -1;
	ssc::SubRow r = ssc::make_sub_row(cdf4);
	return (int)(r.base + (u > r.s0) + (u > r.s1) + (u > r.s2));
}

int64_t ssc_gzip_member_host(const uint8_t* in, uint32_t n, const uint8_t* sample, size_t sample_n, uint8_t* out, size_t cap) {
	if (!in || !out || n < 8) return -1;
	uint64_t hist[256];
	memset(hist, 0, sizeof(hist));
	for (size_t i = 0; i < sample_n; i++) hist[sample[i]]++;
	ssc::GzTables tab;
	const char* err = ssc::gz_build_tables(hist, &tab);
	if (err[0]) { fail(SSC_ERR_INVALID, "gzip tables: %s", err); return -1; }
	return (int64_t)ssc::gz_member_host(&tab, in, n, out, cap);
}

int ssc_issue_floor(ssc_handle* h, int mode, int read_length, int64_t n_pairs, int reps, double* ms_per_launch) {
	if (!h || !ms_per_launch) return fail(SSC_ERR_INVALID, "null argument");
	if (mode < 0 || mode > 1 || read_length < 33 || read_length > 160 || n_pairs < FG_CHUNK || n_pairs > (1 << 22) || reps < 1)
		return fail(SSC_ERR_INVALID, "ssc_issue_floor: mode 0/1, read_length 33..160, n_pairs 32..4194304, reps >= 1");
	CK(cudaSetDevice(h->device));
	const int64_t nTiles = (n_pairs + FG_CHUNK - 1) / FG_CHUNK;
	const uint32_t blobPitch = FG_CHUNK * FG_SLOT;
	DevBuf<uint32_t> scratch; DevBuf<uint8_t> blobs;
	CK(scratch.alloc(1 + (size_t)h->smCount * FG_WORKERS));
	const size_t fileBytes = (size_t)nTiles * blobPitch + 256;
	if (mode == 1) {
		if (2 * fileBytes >= (1ull << 32)) return fail(SSC_ERR_INVALID, "ssc_issue_floor: n_pairs too large for 32-bit blob cursors");
		CK(blobs.alloc(2 * fileBytes));
	}
	cudaStream_t s = h->compute;
	const int grid = (int)std::min<int64_t>((nTiles + FG_WORKERS - 1) / FG_WORKERS, h->smCount);
	CK(ssc::launch_issue_floor(mode, read_length, n_pairs, h->seed, grid, scratch.p, blobs.p, blobPitch, (uint32_t)fileBytes, s));   // warm-up
	CK(cudaEventRecord(h->evStart, s));
	for (int r = 0; r < reps; r++)
		CK(ssc::launch_issue_floor(mode, read_length, n_pairs, h->seed + 1 + (uint64_t)r, grid, scratch.p, blobs.p, blobPitch, (uint32_t)fileBytes, s));
	CK(cudaEventRecord(h->evStop, s));
	CK(cudaEventSynchronize(h->evStop));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, h->evStart, h->evStop));
	*ms_per_launch = (double)ms / reps;
	h->stats.launches += (uint64_t)reps + 1;
	scratch.release(); blobs.release();
	return SSC_OK;
}

int ssc_get_stats(ssc_handle* h, ssc_stats* out) {
	if (!h || !out) return fail(SSC_ERR_INVALID, "null argument");
	*out = h->stats;
	return SSC_OK;
}

int ssc_reset_stats(ssc_handle* h) {
	if (!h) return fail(SSC_ERR_INVALID, "null handle");
	memset(&h->stats, 0, sizeof(h->stats));
	return SSC_OK;
}

}  // extern "C"
