// host_profile.cpp -- .profile parser and the bit-defining table construction
// (Profile::load lib/profile/Profile.cpp:934-1238, normParas(true) :836-932, initCDFs :1367-1434,
//  Matrix::normalize / cumsum lib/matrix/Matrix.h:482-522, normpdf lib/mydefine/MyDefine.cpp:53-56).
// Compiled with -ffp-contract=off: every sum and quotient below must round like the reference.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "host.h"

namespace sschost {

static const double ZERO_FINAL = 2.2204e-16;

// getNextLine, lib/mydefine/MyDefine.cpp:239-251: next line that is neither empty nor a '#' comment
static bool next_line(std::ifstream& ifs, std::string& line, int& lineNum) {
	line = "";
	while (std::getline(ifs, line)) {
		lineNum++;
		if (!line.empty() && line.at(0) != '#') break;
	}
	return !line.empty();
}

// row /= (ZF + left-to-right row sum)   (Matrix::normalize(0))
static void normalize_rows(std::vector<double>& m, int rows, int cols) {
	for (int i = 0; i < rows; i++) {
		double s = 0;
		for (int j = 0; j < cols; j++) s += m[(size_t)i * cols + j];
		for (int j = 0; j < cols; j++) m[(size_t)i * cols + j] /= (ZERO_FINAL + s);
	}
}

static void cumsum_rows(std::vector<double>& m, int rows, int cols) {
	for (int i = 0; i < rows; i++)
		for (int j = 1; j < cols; j++) m[(size_t)i * cols + j] = m[(size_t)i * cols + j] + m[(size_t)i * cols + j - 1];
}

static double normpdf(double x, double mu, double sigma) {
	double PI = 3.1415926;
	return exp(-pow(x - mu, 2) / (2 * pow(sigma, 2))) / (sqrt(2 * PI) * sigma);
}

void ProfileModel::load(const std::string& path, Config& cfg) {
	std::ifstream ifs(path.c_str());
	if (!ifs.is_open()) die(-1, "can not open file " + path);
	std::string line;
	int lineNum = 0;
	const std::string errMsg = "Error: malformed model file " + path + " @line ";
	int binCount = -1, kmer = -1, readLength = -1;
	bases = "";
	while (next_line(ifs, line, lineNum)) {
		std::vector<std::string> f = split(line, ':');
		if (f.size() != 2) die(1, errMsg + std::to_string(lineNum) + "\n" + line);
		std::string key = trim(f[0]);
		if (key == "bases") { bases = trim(f[1]); if (bases.empty()) die(1, errMsg + std::to_string(lineNum) + "\n" + line); }
		else if (key == "binCount") { binCount = atoi(trim(f[1]).c_str()); if (binCount <= 0) die(1, errMsg + std::to_string(lineNum) + "\n" + line); }
		else if (key == "kmer") { kmer = atoi(trim(f[1]).c_str()); if (kmer <= 0) die(1, errMsg + std::to_string(lineNum) + "\n" + line); }
		else if (key == "readLength") { readLength = atoi(trim(f[1]).c_str()); if (readLength <= 0) die(1, errMsg + std::to_string(lineNum) + "\n" + line); }
		else die(1, errMsg + std::to_string(lineNum) + "\n" + line);
		if (!bases.empty() && binCount > 0 && kmer > 0 && readLength > 0) break;
	}
	if (bases.empty() || binCount <= 0 || kmer <= 0 || readLength <= 0) die(1, "Error: malformed model file " + path);
	// the profile header overrides the configuration (Profile.cpp:1000-1003); bins are capped by the read length (:184-188)
	// (the reference clamps bins to the read length in Profile::init but keeps parsing binCount rows into the smaller
	//  matrices, Profile.cpp:1000-1108 -- heap corruption; we refuse such a profile instead)
	if (binCount > readLength) die(1, "Error: malformed model file " + path + ": binCount larger than readLength");
	cfg.str["bases"] = bases; cfg.num["kmer"] = kmer; cfg.num["bins"] = binCount; cfg.num["readLength"] = readLength;
	N = (int)bases.length(); K = kmer; B = binCount; RL = readLength;
	rows = 0;
	{ long pw = N; for (int p = 1; p <= K; p++) { rows += (int)pw; pw *= N; } }

	auto base_index = [&](char c) { for (int i = 0; i < N; i++) if (bases[i] == c) return i; return -1; };
	// row of a k-mer string in Profile::initKmers order (Profile.cpp:70-124)
	auto kmer_row = [&](const std::string& s) -> int {
		if ((int)s.length() != K) return -1;
		int pad = 0;
		while (pad < K && s[pad] == 'X') pad++;
		int valid = K - pad;
		if (valid < 1) return -1;
		int offset = 0; long pw = N;
		for (int q = 1; q < valid; q++) { offset += (int)pw; pw *= N; }
		int v = 0;
		for (int i = pad; i < K; i++) { int b = base_index(s[i]); if (b < 0) return -1; v = v * N + b; }
		return offset + v;
	};

	std::vector<double> insFreqs(1, 0.0), delFreqs(1, 0.0);
	sub1.assign((size_t)rows * B * N, 0.0);
	sub2.assign((size_t)rows * B * N, 0.0);
	qual.assign((size_t)N * N * B * Q, 0.0);
	std::vector<int> lastBase(rows, 0);
	for (int r = 0; r < rows; r++) lastBase[r] = r % N;   // getIndexOfBase(kmers[i][kmer-1]): group offsets are multiples of N
	for (int i = 0; i < 101; i++) gcMeans[i] = 0;   // (the reference leaves them uninitialised; every shipped profile sets all 101)
	int loaded = 0;
	const std::string trunc = "Error: malformed profile file " + path;
	while (next_line(ifs, line, lineNum)) {
		if (line == "[Insert Rate]") {
			if (!next_line(ifs, line, lineNum)) die(1, trunc);
			insertRate = atof(trim(line).c_str());
			loaded++;
		} else if (line == "[Insert Frequency]" || line == "[Deletion Frequency]") {
			bool isIns = line == "[Insert Frequency]";
			if (!next_line(ifs, line, lineNum)) die(1, trunc);
			std::vector<std::string> f = split(line, '\t');
			if (f.size() < 1) die(1, errMsg + std::to_string(lineNum) + "\n" + line);
			std::vector<double>& v = isIns ? insFreqs : delFreqs;
			v.resize(f.size());
			for (size_t j = 0; j < f.size(); j++) v[j] = atof(trim(f[j]).c_str());
			loaded++;
		} else if (line == "[Deletion Rate]") {
			if (!next_line(ifs, line, lineNum)) die(1, trunc);
			delRate = atof(trim(line).c_str());
			loaded++;
		} else if (line == "[Substitution Probs]") {
			for (int i = 0; i < rows; i++) {
				if (!next_line(ifs, line, lineNum)) die(1, trunc);
				std::vector<std::string> f = split(line, ':');
				if (f.size() != 2 || trim(f[0]) != "kmer") die(1, errMsg + std::to_string(lineNum) + "\n" + line);
				std::string ks = trim(f[1]);
				int row = kmer_row(ks);
				if (row == -1) die(1, "Error: unrecognized kmer @line " + std::to_string(lineNum) + " in profile file " + path + "\n" + line);
				for (int j = 0; j < B * 2; j++) {
					if (!next_line(ifs, line, lineNum)) die(1, trunc);
					std::vector<std::string> v = split(line, '\t');
					if ((int)v.size() != N) die(1, errMsg + std::to_string(lineNum) + "\n" + line);
					for (int k = 0; k < N; k++) {
						double p = atof(trim(v[k]).c_str());
						if (j < B) sub1[((size_t)row * B + j) * N + k] = p;
						else sub2[((size_t)row * B + (j - B)) * N + k] = p;
					}
				}
			}
			loaded++;
		} else if (line == "[Base Quality Distribution]") {
			for (int i = 0; i < N * N; i++) {
				if (!next_line(ifs, line, lineNum)) die(1, trunc);
				std::vector<std::string> f = split(line, ':');
				if (f.size() != 2 || trim(f[0]) != "basePairIndx") die(1, errMsg + std::to_string(lineNum) + "\n" + line);
				int bp = atoi(trim(f[1]).c_str());
				if (bp < 0 || bp > N * N - 1) die(1, "Error: unrecognized basePairIndx @line " + std::to_string(lineNum) + " in profile file " + path + "\n" + line);
				for (int j = 0; j < B; j++) {
					if (!next_line(ifs, line, lineNum)) die(1, trunc);
					std::vector<std::string> v = split(line, '\t');
					if ((int)v.size() != Q) die(1, errMsg + std::to_string(lineNum) + "\n" + line);
					for (int k = 0; k < Q; k++) qual[((size_t)bp * B + j) * Q + k] = atof(trim(v[k]).c_str());
				}
			}
			loaded++;
		} else if (line == "[Insert Size Standard Deviation]") {
			if (!next_line(ifs, line, lineNum)) die(1, "Error: malformed model file " + path);
			stdISize = atof(trim(line).c_str());
			loaded++;
		} else if (line == "[Log Ratio Mean Value]") {
			for (int j = 0; j < 101; j++) {
				if (!next_line(ifs, line, lineNum)) die(1, "Error: malformed model file " + path);
				std::vector<std::string> f = split(line, '\t');
				if (f.size() != 2) die(1, errMsg + std::to_string(lineNum) + "\n" + line);
				int gc = atoi(f[0].c_str());
				if (gc < 0 || gc > 100) die(1, errMsg + std::to_string(lineNum) + "\n" + line);
				gcMeans[gc] = atof(f[1].c_str());
			}
			loaded++;
		} else if (line == "[Log Ratio Standard Deviation]") {
			if (!next_line(ifs, line, lineNum)) die(1, "Error: malformed model file " + path);
			gcStd = atof(trim(line).c_str());
			loaded++;
		}
	}
	if (loaded < 9) die(1, "Error: corrupted model file " + path + ", failed to load some parameters!");
	std::cerr << "profile was loaded from file " << path << std::endl;

	// ---- normParas(true), Profile.cpp:836-932
	for (int t = 0; t < 2; t++) {
		std::vector<double>& m = t ? sub2 : sub1;
		for (int r = 0; r < rows; r++) {
			std::vector<double> row(m.begin() + (size_t)r * B * N, m.begin() + (size_t)(r + 1) * B * N);
			normalize_rows(row, B, N);
			for (int j = 0; j < B; j++) {
				double s = 0;
				for (int k = 0; k < N; k++) s += row[(size_t)j * N + k];
				if (s < ZERO_FINAL) row[(size_t)j * N + lastBase[r]] = 1;   // Profile.cpp:848-860
			}
			std::copy(row.begin(), row.end(), m.begin() + (size_t)r * B * N);
		}
	}
	normalize_rows(qual, N * N * B, Q);                    // Profile.cpp:864-866
	const bool paired = cfg.paired();
	std::vector<double> isizeDist;
	if (paired && stdISize > 0) {                          // Profile.cpp:912-930
		int mean = cfg.num["insertSize"] + 1;
		int intervalLen = (int)(6 * stdISize);
		int mn = std::max(mean - intervalLen / 2, RL);
		int mx = 2 * mean - mn;
		minIS = mn;
		int cnt = mx - mn + 1;
		if (cnt < 1) die(1, "Error: empty insert size table (insertSize smaller than read length?)");
		isizeDist.resize(cnt);
		for (int i = 0; i < cnt; i++) isizeDist[i] = normpdf((double)(mn + i), (double)mean, stdISize);
		normalize_rows(isizeDist, 1, cnt);
	}
	// ---- initCDFs, Profile.cpp:1367-1434
	insCdf = insFreqs; cumsum_rows(insCdf, 1, (int)insCdf.size());
	delCdf = delFreqs; cumsum_rows(delCdf, 1, (int)delCdf.size());
	normalize_rows(qual, N * N * B, Q);                    // second normalisation, :1397
	cumsum_rows(qual, N * N * B, Q);
	isizeCdf = isizeDist;
	if (!isizeCdf.empty()) cumsum_rows(isizeCdf, 1, (int)isizeCdf.size());
	cumsum_rows(sub1, rows * B, N);
	useCdf2 = paired && stdISize > 0;                      // :1420-1430
	if (useCdf2) cumsum_rows(sub2, rows * B, N); else sub2.clear();
}

void ProfileModel::fill(ssc_profile_tables* t, const Config& cfg) const {
	memset(t, 0, sizeof(*t));
	t->n_bases = N; t->kmer = K; t->bins = B; t->n_qual = Q; t->min_qual = minQ; t->read_length = RL;
	t->paired = cfg.paired() ? 1 : 0; t->use_cdf2 = useCdf2 ? 1 : 0;
	t->fixed_insert_size = cfg.num.at("insertSize"); t->min_insert_size = minIS;
	t->n_isize = (int)isizeCdf.size(); t->n_ins = (int)insCdf.size(); t->n_del = (int)delCdf.size();
	t->n_kmer_rows = rows; t->insert_rate = insertRate; t->del_rate = delRate;
	strncpy(t->bases, bases.c_str(), 8);
	t->isize_cdf = isizeCdf.empty() ? nullptr : isizeCdf.data();
	t->ins_cdf = insCdf.data(); t->del_cdf = delCdf.data();
	t->subs_cdf1 = sub1.data(); t->subs_cdf2 = useCdf2 ? sub2.data() : nullptr; t->quality_cdf = qual.data();
}

void ProfileModel::seed_gc(uint64_t seed) {
	gcEng.clear(); gcDist.clear();
	for (unsigned l = 0; l < 101; l++) {                   // Profile.cpp:1409-1415, seeds pinned to seed + l
		gcEng.push_back(std::default_random_engine((unsigned)(seed + l)));
		gcDist.push_back(std::normal_distribution<double>(gcMeans[l], gcStd));
	}
}

double ProfileModel::gc_factor(int gc) {                   // Profile::getGCFactor, Profile.cpp:1507-1517
	if (gc < 0 || gc > 100) return 0;
	double v = gcDist[gc](gcEng[gc]);
	while (v < 0) v = gcDist[gc](gcEng[gc]);
	return v;
}

}  // namespace sschost
