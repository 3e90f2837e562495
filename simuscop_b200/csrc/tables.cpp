// tables.cpp -- see tables.h.  Compiled with -ffp-contract=off: the FP64 expressions below
// must round exactly like the reference's x86-64 build (no FMA).
#include "tables.h"

#include <cstring>
#include <algorithm>
#include <functional>

namespace ssc {

static const double ZERO_FINAL = 2.2204e-16;  // lib/mydefine/MyDefine.cpp:20

double draw_real(uint32_t u, double start, double end) {
	// ThreadPool::randomDouble, lib/threadpool/ThreadPool.cpp:203-207 with min=0, max=2^32-1
	volatile double number = (double)u;
	volatile double frac = number / 4294967296.0;
	volatile double scaled = (end - start) * frac;
	return start + scaled;
}

// number of u in [0, 2^32) for which pred(u) holds; pred must be true on a prefix of the range
static uint64_t count_prefix(const std::function<bool(uint32_t)>& pred) {
	if (!pred(0)) return 0;
	if (pred(0xFFFFFFFFu)) return 1ull << 32;
	uint64_t lo = 0, hi = 0xFFFFFFFFull;  // pred(lo) true, pred(hi) false
	while (hi - lo > 1) {
		uint64_t mid = (lo + hi) >> 1;
		if (pred((uint32_t)mid)) lo = mid; else hi = mid;
	}
	return lo + 1;
}

uint64_t count_le(double c) {
	return count_prefix([c](uint32_t u) { return draw_real(u, ZERO_FINAL, 1) <= c; });
}

CompressedCdf compress_cdf(const double* cdf, int ac) {
	CompressedCdf out;
	uint64_t M = 0;  // draws below M are already claimed by an earlier symbol
	for (int k = 0; k < ac; k++) {
		uint64_t C = count_le(cdf[k]);
		if (C > M) {
			out.T.push_back((uint32_t)(C - 1));
			out.sym.push_back((uint16_t)k);
			M = C;
		}
	}
	if (M < (1ull << 32)) {  // "return ac-1" fall-through of randIndx
		if (!out.sym.empty() && out.sym.back() == (uint16_t)(ac - 1)) out.T.back() = 0xFFFFFFFFu;
		else { out.T.push_back(0xFFFFFFFFu); out.sym.push_back((uint16_t)(ac - 1)); }
	}
	return out;
}

SubRow make_sub_row(const double* cdf4) {
	// call = #{k<3 : u >= M_k},  M_k = max_{j<=k} count_le(cdf[j])   (first k with r <= cdf[k], else 3)
	uint64_t M[3];
	uint64_t run = 0;
	for (int k = 0; k < 3; k++) {
		uint64_t C = count_le(cdf4[k]);
		if (C > run) run = C;
		M[k] = run;
	}
	SubRow r;
	uint32_t s[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
	uint32_t base = 0;
	int n = 0;
	for (int k = 0; k < 3; k++) {
		if (M[k] == 0) base++;                 // u >= 0 always
		else s[n++] = (uint32_t)(M[k] - 1);    // u >= M  <=>  u > M-1
	}
	r.s0 = s[0]; r.s1 = s[1]; r.s2 = s[2]; r.base = base;
	return r;
}

const char* build_tables(const ssc_profile_tables* t, DeviceTablesHost* o) {
	if (!t) return "null profile";
	if (t->n_bases != 4) return "unsupported alphabet: n_bases must be 4";
	if (t->kmer < 1 || t->kmer > 8) return "unsupported kmer (1..8)";
	if (t->bins < 1 || t->read_length < 1) return "bad bins/read_length";
	if (t->n_qual < 1 || t->n_qual > 256) return "bad n_qual";
	if (!t->subs_cdf1 || !t->quality_cdf || !t->ins_cdf || !t->del_cdf) return "missing table pointer";
	if (t->n_isize > 0 && !t->isize_cdf) return "missing isize_cdf";
	if (t->use_cdf2 && !t->subs_cdf2) return "missing subs_cdf2";
	int rows = 0, pw = 4;
	for (int p = 1; p <= t->kmer; p++) { rows += pw; pw *= 4; }
	if (rows != t->n_kmer_rows) return "n_kmer_rows does not match kmer";
	o->N = 4; o->K = t->kmer; o->B = t->bins; o->Q = t->n_qual; o->minQ = t->min_qual; o->RL = t->read_length;
	o->paired = t->paired ? 1 : 0; o->useCdf2 = t->use_cdf2 ? 1 : 0;
	o->fixedInsert = t->fixed_insert_size; o->minIS = t->min_insert_size; o->nRows = rows;

	// alphabet
	for (int i = 0; i < 256; i++) o->asciiCode[i] = 4;
	bool seen[4] = {false, false, false, false};
	const char* canon = "ACGT";
	for (int i = 0; i < 4; i++) {
		char c = t->bases[i];
		const char* p = c ? strchr(canon, c) : nullptr;
		if (!p) return "unsupported alphabet: bases must be a permutation of ACGT";
		if (seen[p - canon]) return "unsupported alphabet: repeated base";
		seen[p - canon] = true;
		o->baseChar[i] = c;
		o->asciiCode[(unsigned char)c] = (int8_t)i;
		o->asciiCode[(unsigned char)(c | 0x20)] = (int8_t)i;  // haplotypes are upper-cased (Segment.cpp:448-458)
	}
	o->compLut = 0;
	for (int i = 0; i < 4; i++) {
		char c = t->bases[i];
		char cc = (c == 'A') ? 'T' : (c == 'T') ? 'A' : (c == 'C') ? 'G' : 'C';  // Segment::getComplementSeq
		o->compLut |= (uint8_t)(o->asciiCode[(unsigned char)cc] << (2 * i));
	}

	// scalar event tests (Profile::getIndelSeq, lib/profile/Profile.cpp:1560-1572)
	double ir = t->insert_rate;
	uint64_t ci = count_prefix([ir](uint32_t u) { return draw_real(u, 0, 1) <= ir; });
	o->insEnable = ci > 0; o->insT = ci > 0 ? (uint32_t)(ci - 1) : 0;
	volatile double one_minus = 1 - t->insert_rate;
	double d = t->del_rate / one_minus;
	uint64_t cd = count_prefix([d](uint32_t u) { return draw_real(u, 0, 1) < d; });
	o->delEnable = cd > 0; o->delT = cd > 0 ? (uint32_t)(cd - 1) : 0;

	if (t->n_isize > 0) o->isize = compress_cdf(t->isize_cdf, t->n_isize);
	else { o->isize.T.clear(); o->isize.sym.clear(); }
	o->insLen = compress_cdf(t->ins_cdf, t->n_ins);
	o->delLen = compress_cdf(t->del_cdf, t->n_del);

	size_t nsub = (size_t)rows * t->bins;
	o->sub.resize(nsub * (o->useCdf2 ? 2 : 1));
	for (size_t i = 0; i < nsub; i++) o->sub[i] = make_sub_row(t->subs_cdf1 + i * 4);
	if (o->useCdf2) for (size_t i = 0; i < nsub; i++) o->sub[nsub + i] = make_sub_row(t->subs_cdf2 + i * 4);

	size_t nq = (size_t)16 * t->bins;
	std::vector<CompressedCdf> qrows(nq);
	size_t maxLen = 1;
	for (size_t i = 0; i < nq; i++) {
		qrows[i] = compress_cdf(t->quality_cdf + i * t->n_qual, t->n_qual);
		if (qrows[i].T.size() > maxLen) maxLen = qrows[i].T.size();
	}
	int pitch = 1;
	while ((size_t)pitch < maxLen) pitch <<= 1;
	o->qualPitch = pitch;
	o->maxQualRow = (int)maxLen;
	o->qualT.assign(nq * pitch, 0xFFFFFFFFu);
	o->qualSym.assign(nq * pitch, 0);
	for (size_t i = 0; i < nq; i++) {
		size_t n = qrows[i].T.size();
		for (int j = 0; j < pitch; j++) {
			size_t s = (size_t)j < n ? (size_t)j : n - 1;
			o->qualT[i * pitch + j] = (size_t)j < n ? qrows[i].T[j] : 0xFFFFFFFFu;
			o->qualSym[i * pitch + j] = (uint8_t)(t->min_qual + qrows[i].sym[s]);
		}
	}
	// diagonal rows (ref == call)
	size_t dmax = 1;
	for (int b = 0; b < 4; b++)
		for (int j = 0; j < t->bins; j++) dmax = std::max(dmax, qrows[(size_t)(b * 4 + b) * t->bins + j].T.size());
	o->diagPitch = (int)((dmax + 3) / 4 * 4);
	o->qualDiagT.assign((size_t)4 * t->bins * o->diagPitch, 0xFFFFFFFFu);
	o->qualDiagSym.assign((size_t)4 * t->bins * o->diagPitch, 0);
	for (int b = 0; b < 4; b++)
		for (int j = 0; j < t->bins; j++) {
			const CompressedCdf& r = qrows[(size_t)(b * 4 + b) * t->bins + j];
			size_t n = r.T.size(), base = ((size_t)b * t->bins + j) * o->diagPitch;
			for (int k = 0; k < o->diagPitch; k++) {
				size_t sidx = (size_t)k < n ? (size_t)k : n - 1;
				o->qualDiagT[base + k] = (size_t)k < n ? r.T[k] : 0xFFFFFFFFu;
				o->qualDiagSym[base + k] = (uint8_t)(t->min_qual + r.sym[sidx]);
			}
		}
	return "";
}

}  // namespace ssc
