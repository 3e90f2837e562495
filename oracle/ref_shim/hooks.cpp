/*
 * hooks.cpp -- TEST INFRASTRUCTURE (oracle side). Not part of the product.
 *
 * Cursor for the Philox-instrumented reference build plus the plan dump.
 * Compiled against the reference's own headers; linked with the reference's
 * objects (scratch-patched by oracle/build_ref.py) and philox_pool.cpp.
 *
 * Environment:
 *   SIMUSCOP_SEED        64-bit seed (default 1): Philox key, srand() seed and
 *                        GC-engine seeds (seed + l).
 *   SIMUSCOP_DUMP_PLAN   if set, prefix of the SSCPLAN1 files written per
 *                        output sample: <prefix>.<sample>.plan
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include <random>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <cmath>
#include <cassert>
#include <pthread.h>
#include <sys/time.h>
#include <unistd.h>

#define private public
#define protected public
#include "MyDefine.h"
#include "Segment.h"
#undef private
#undef protected

#include "ssc_hooks.h"
#include "../ssc_oracle_philox.h"

namespace {

struct Cursor {
	bool init;
	uint64_t seed;
	int paired;
	uint64_t pairBase, pairBaseNext;
	int64_t ordinal;
	int attempt;
	bool newPair;
	int mate;
	int refpos, outpos;
	int slot;
	int insBase;
	Cursor() : init(false), seed(1), paired(0), pairBase(0), pairBaseNext(0), ordinal(-1), attempt(0),
		newPair(true), mate(0), refpos(0), outpos(0), slot(SSC_SLOT_NONE), insBase(0) {}
};

Cursor cur;

void ensureInit() {
	if (cur.init) return;
	cur.init = true;
	const char* s = getenv("SIMUSCOP_SEED");
	cur.seed = s ? strtoull(s, NULL, 10) : 1ULL;
	cur.paired = config.isPairedEnd() ? 1 : 0;
}

/* ---------- plan dump ---------- */
FILE* planFp = NULL;
int sampleIndex = 0;
std::string lastUnit;

void putRec(int32_t tag, const std::string& payload) {
	int64_t n = (int64_t)payload.size();
	fwrite(&tag, 4, 1, planFp);
	fwrite(&n, 8, 1, planFp);
	fwrite(payload.data(), 1, payload.size(), planFp);
}
template <class T> void app(std::string& s, T v) { s.append((const char*)&v, sizeof(T)); }
void appDoubles(std::string& s, const double* p, size_t n) { if (n) s.append((const char*)p, n*sizeof(double)); }

void dumpProfile() {
	std::string p;
	std::string bases = config.getStringPara("bases");
	int N = bases.length();
	int K = config.getIntPara("kmer");
	int B = config.getIntPara("bins");
	int Q = profile.maxBaseQuality-profile.minBaseQuality+1;
	int useCdf2 = (profile.subsCdf2[0].getEntrance() != NULL) ? 1 : 0;
	int nIS = (profile.iSizeAlphabet.getEntrance() != NULL) ? profile.iSizeAlphabet.getCOLS() : 0;
	int minIS = nIS ? profile.iSizeAlphabet.get(0, 0) : 0;
	app<int32_t>(p, N); app<int32_t>(p, K); app<int32_t>(p, B); app<int32_t>(p, Q);
	app<int32_t>(p, profile.minBaseQuality);
	app<int32_t>(p, config.getIntPara("readLength"));
	app<int32_t>(p, cur.paired);
	app<int32_t>(p, useCdf2);
	app<int32_t>(p, config.getIntPara("insertSize"));
	app<int32_t>(p, minIS);
	app<int32_t>(p, nIS);
	app<int32_t>(p, profile.insCdf.getCOLS());
	app<int32_t>(p, profile.delCdf.getCOLS());
	app<int32_t>(p, profile.kmerCount);
	app<int32_t>(p, config.getIntPara("ploidy"));
	app<int32_t>(p, 0);
	app<double>(p, profile.insertRate);
	app<double>(p, profile.delRate);
	char b8[8]; memset(b8, 0, 8); strncpy(b8, bases.c_str(), 8);
	p.append(b8, 8);
	if (nIS) appDoubles(p, profile.iSizeCdf.getEntrance(), nIS);
	appDoubles(p, profile.insCdf.getEntrance(), profile.insCdf.getCOLS());
	appDoubles(p, profile.delCdf.getEntrance(), profile.delCdf.getCOLS());
	for (int i = 0; i < profile.kmerCount; i++) appDoubles(p, profile.subsCdf1[i].getEntrance(), (size_t)B*N);
	if (useCdf2) for (int i = 0; i < profile.kmerCount; i++) appDoubles(p, profile.subsCdf2[i].getEntrance(), (size_t)B*N);
	for (int i = 0; i < N*N; i++) appDoubles(p, profile.qualityCdf[i].getEntrance(), (size_t)B*Q);
	putRec(1, p);
}

void openPlanIfNeeded() {
	if (planFp) return;
	const char* prefix = getenv("SIMUSCOP_DUMP_PLAN");
	if (!prefix) return;
	char fn[4096];
	snprintf(fn, sizeof(fn), "%s.%d.plan", prefix, sampleIndex);
	planFp = fopen(fn, "wb");
	if (!planFp) { fprintf(stderr, "ssc hooks: cannot open %s\n", fn); exit(3); }
	fwrite("SSCPLAN1", 1, 8, planFp);
	dumpProfile();
	lastUnit = "";
}

} // namespace

void ssc_hook_dump_chr(std::vector<Segment>& chrSegs, const std::string& popu, const std::string& chr) {
	ensureInit();
	if (!getenv("SIMUSCOP_DUMP_PLAN")) return;
	openPlanIfNeeded();
	int ploidy = config.getIntPara("ploidy");
	{
		std::string u;
		app<int32_t>(u, (int32_t)popu.size()); app<int32_t>(u, (int32_t)chr.size());
		u += popu; u += chr;
		putRec(2, u);
	}
	for (size_t k = 0; k < chrSegs.size(); k++) {
		Segment& seg = chrSegs[k];
		std::string s;
		char** seqs = seg.getSegSequences();
		int nb = (int)seg.fragStartPos.size();
		int64_t segsize = 0;
		if (seqs != NULL && seg.getCN() > 0) segsize = seg.getSeqSize()/seg.getCN();
		app<int32_t>(s, seg.getSegIndx()); app<int32_t>(s, seg.getCN());
		app<int64_t>(s, seg.getSegStartPos()); app<int64_t>(s, seg.getSegEndPos());
		app<int64_t>(s, segsize);
		app<int64_t>(s, seg.getReadCount());
		app<int32_t>(s, nb); app<int32_t>(s, ploidy);
		for (int h = 0; h < ploidy; h++) {
			int64_t L = (seqs && seqs[h]) ? (int64_t)strlen(seqs[h]) : 0;
			app<int64_t>(s, L);
		}
		for (int h = 0; h < ploidy; h++) if (seqs && seqs[h]) s.append(seqs[h], strlen(seqs[h]));
		for (int i = 0; i < nb; i++) app<int64_t>(s, seg.fragStartPos[i]);
		for (int i = 0; i < nb; i++) app<int64_t>(s, seg.fragEndPos[i]);
		for (int i = 0; i < nb; i++) app<int32_t>(s, seg.hapIndxs[i]);
		for (int i = 0; i < nb; i++) app<int32_t>(s, (i < (int)seg.fragRCs.size()) ? seg.fragRCs[i] : 0);
		putRec(3, s);
	}
}

void ssc_hook_sample_end() {
	if (planFp) {
		putRec(9, std::string());
		fclose(planFp);
		planFp = NULL;
	}
	sampleIndex++;
	/* pair IDs restart at 0 for every output sample (one plan per sample) */
	cur.pairBase = cur.pairBaseNext = 0;
}

void ssc_hook_bin(int readCount) {
	ensureInit();
	cur.pairBase = cur.pairBaseNext;
	if (readCount > 0) cur.pairBaseNext += cur.paired ? (uint64_t)((readCount+1)/2) : (uint64_t)readCount;
	cur.ordinal = -1;
	cur.newPair = true;
	cur.slot = SSC_SLOT_NONE;
}

void ssc_hook_attempt() {
	if (cur.newPair) { cur.ordinal++; cur.attempt = 0; cur.newPair = false; }
	else cur.attempt++;
	cur.slot = SSC_SLOT_POS;
}

void ssc_hook_frag_ok() { cur.newPair = true; }

void ssc_hook_slot(int slot) { cur.slot = slot; if (slot == SSC_SLOT_INSLEN) cur.insBase = 0; }

void ssc_hook_mate(int isRead1) { cur.mate = isRead1 ? 0 : 1; cur.slot = SSC_SLOT_NONE; }

void ssc_hook_refpos(int j) { cur.refpos = j; cur.slot = SSC_SLOT_P; }

void ssc_hook_outpos(int j) { cur.outpos = j; cur.slot = SSC_SLOT_NONE; }

unsigned ssc_hook_gc_seed(unsigned l) { ensureInit(); return (unsigned)(cur.seed + l); }

unsigned ssc_hook_plan_seed() { ensureInit(); return (unsigned)cur.seed; }

unsigned int ssc_next_u32(bool integerDraw) {
	ensureInit();
	uint64_t pair = cur.pairBase + (uint64_t)cur.ordinal;
	uint32_t w[4];
	int word = -1;
	switch (cur.slot) {
	case SSC_SLOT_POS:
		if (!integerDraw) goto bad;
		ssco_block(cur.seed, pair, 0, SSCO_STREAM_FRAG, 0, (uint32_t)cur.attempt, w); word = 0;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_ISIZE:
		if (integerDraw) goto bad;
		ssco_block(cur.seed, pair, 0, SSCO_STREAM_FRAG, 0, (uint32_t)cur.attempt, w); word = 1;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_STRAND:
		if (!integerDraw) goto bad;
		ssco_block(cur.seed, pair, 0, SSCO_STREAM_FRAG, 0, (uint32_t)cur.attempt, w); word = 2;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_P:
		if (integerDraw) goto bad;
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_CYCLE, 0, (uint32_t)cur.refpos, w); word = 0;
		cur.slot = SSC_SLOT_P2; break;
	case SSC_SLOT_P2:
		if (integerDraw) goto bad;
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_CYCLE, 0, (uint32_t)cur.refpos, w); word = 1;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_SUB:
		if (integerDraw) goto bad;
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_CYCLE, 0, (uint32_t)cur.outpos, w); word = 2;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_QUAL:
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_CYCLE, 0, (uint32_t)cur.outpos, w); word = 3;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_INSLEN:
		if (integerDraw) goto bad;
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_LEN, 0, (uint32_t)cur.refpos, w); word = 0;
		cur.slot = SSC_SLOT_INSBASE; cur.insBase = 0; break;
	case SSC_SLOT_DELLEN:
		if (integerDraw) goto bad;
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_LEN, 0, (uint32_t)cur.refpos, w); word = 1;
		cur.slot = SSC_SLOT_NONE; break;
	case SSC_SLOT_INSBASE:
		if (!integerDraw) goto bad;
		ssco_block(cur.seed, pair, cur.mate, SSCO_STREAM_INSBASE, cur.insBase/4, (uint32_t)cur.refpos, w);
		word = cur.insBase%4; cur.insBase++; break;
	default:
		goto bad;
	}
	return w[word];
bad:
	fprintf(stderr, "ssc hooks: unexpected %s draw in slot %d (pair %llu mate %d refpos %d outpos %d)\n",
		integerDraw ? "integer" : "real", cur.slot, (unsigned long long)pair, cur.mate, cur.refpos, cur.outpos);
	abort();
}
