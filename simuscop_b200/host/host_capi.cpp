// host_capi.cpp -- C ABI of include/simuscop_host.h and the file sink of the drop-in run.
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/simuscop_host.h"
#include "host.h"

struct ssh_job {
	sschost::Job job;
};

namespace {
struct FileSink {
	int fd1 = -1, fd2 = -1;
};

int write_all(int fd, const char* p, size_t n) {
	while (n > 0) {
		ssize_t w = write(fd, p, n);
		if (w < 0) return 1;
		p += w; n -= (size_t)w;
	}
	return 0;
}

// SeqWriter::write(char*, char*), lib/seqwriter/SeqWriter.cpp:49-54: both files advance together
int file_sink(void* user, const char* b1, size_t l1, const char* b2, size_t l2, int64_t, int64_t) {
	FileSink* s = (FileSink*)user;
	if (write_all(s->fd1, b1, l1)) return 1;
	if (s->fd2 >= 0 && write_all(s->fd2, b2, l2)) return 1;
	return 0;
}
}  // namespace

extern "C" {

int ssh_open(const char* config_path, uint64_t seed, ssh_job** out) {
	if (!config_path || !out) return SSC_ERR_INVALID;
	ssh_job* j = new ssh_job();
	j->job.seed = seed;
	j->job.open(config_path);
	*out = j;
	return SSC_OK;
}

int ssh_close(ssh_job* job) { delete job; return SSC_OK; }
int ssh_num_samples(ssh_job* job) { return (int)job->job.samples.size(); }
const char* ssh_sample_stem(ssh_job* job, int s) { return job->job.samples.at(s).stem.c_str(); }
int ssh_paired(ssh_job* job) { return job->job.cfg.paired() ? 1 : 0; }
int ssh_read_length(ssh_job* job) { return job->job.prof.RL; }
const char* ssh_output_dir(ssh_job* job) { return job->job.cfg.str["output"].c_str(); }

int ssh_prepare_sample(ssh_job* job, int s, ssc_handle* dev, const char* dump_path, int64_t* planned, int64_t* emitted) {
	if (!job || s < 0 || s >= (int)job->job.samples.size()) return SSC_ERR_INVALID;
	return job->job.prepare_sample(s, dev, dump_path ? dump_path : "", planned, emitted);
}

int ssh_run(ssh_job* job, int device) {
	if (getenv("SIMUSCOP_PLAN_ONLY")) {
		// host logic only (no GPU): write the SSCPLAN1 dumps of every sample and stop
		const char* prefix = getenv("SIMUSCOP_DUMP_PLAN");
		if (!prefix) { std::cerr << "Error: SIMUSCOP_PLAN_ONLY needs SIMUSCOP_DUMP_PLAN" << std::endl; return 1; }
		for (int s = 0; s < (int)job->job.samples.size(); s++) {
			int64_t a, b;
			int r = job->job.prepare_sample(s, nullptr, std::string(prefix) == "none" ? std::string() : std::string(prefix) + "." + std::to_string(s) + ".plan", &a, &b);
			if (r) return r;
		}
		return 0;
	}
	// devices: SIMUSCOP_DEVICES="0,1,2,..." shards every sample by pair-ID range over the listed GPUs
	std::vector<int> devIds;
	if (const char* dl = getenv("SIMUSCOP_DEVICES")) {
		for (const std::string& t : sschost::split(dl, ',')) if (!sschost::trim(t).empty()) devIds.push_back(atoi(t.c_str()));
	}
	if (devIds.empty()) devIds.push_back(device);
	std::vector<ssc_handle*> devs;
	int rc = 0;
	for (int id : devIds) {
		ssc_handle* h = nullptr;
		rc = ssc_create(id, &h);
		if (rc) { std::cerr << "Error: " << ssc_last_error() << std::endl; for (auto* d : devs) ssc_destroy(d); return rc; }
		devs.push_back(h);
	}
	const char* dumpPrefix = getenv("SIMUSCOP_DUMP_PLAN");
	if (const char* bp = getenv("SIMUSCOP_BATCH_PAIRS")) for (auto* d : devs) ssc_set_option(d, "batch_pairs", atoll(bp));
	// SIMUSCOP_GZIP=1: the same FASTQ, compressed on the GPU, written as <name>.fq.gz (concatenated gzip members)
	const char* gzEnv = getenv("SIMUSCOP_GZIP");
	const bool gz = gzEnv && atoi(gzEnv) != 0;
	if (gz) for (auto* d : devs) ssc_set_option(d, "gzip", 1);
	sschost::Job& J = job->job;
	const int G = (int)devs.size();
	for (int s = 0; s < (int)J.samples.size() && !rc; s++) {
		const std::string prefix = J.cfg.str["output"] + "/" + J.samples[s].stem;
		const bool paired = J.cfg.paired();
		const std::string ext = gz ? ".fq.gz" : ".fq";
		const std::string f1 = paired ? prefix + "_1" + ext : prefix + ext;
		const std::string f2 = paired ? prefix + "_2" + ext : "";
		std::string dump;
		if (dumpPrefix) dump = std::string(dumpPrefix) + "." + std::to_string(s) + ".plan";
		int64_t planned = 0, emitted = 0;
		rc = J.prepare_sample_multi(s, devs, dump, &planned, &emitted);
		if (rc) { std::cerr << "Error: " << ssc_last_error() << std::endl; break; }
		// shard g writes <file>.part<g> (g > 0) or the final file (g == 0); parts are appended in rank order afterwards
		std::vector<FileSink> sinks(G);
		std::vector<int> rcs(G, 0);
		std::vector<std::string> errs(G);
		auto part = [&](const std::string& f, int g) { return g == 0 ? f : f + ".part" + std::to_string(g); };
		for (int g = 0; g < G; g++) {
			sinks[g].fd1 = open(part(f1, g).c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
			if (sinks[g].fd1 < 0) sschost::die(-1, "Error: can not open fastq file to save results:\n" + part(f1, g));
			if (paired) {
				sinks[g].fd2 = open(part(f2, g).c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
				if (sinks[g].fd2 < 0) sschost::die(-1, "Error: can not open fastq file to save results:\n" + part(f2, g));
			}
		}
		auto work = [&](int g) {
			const int64_t base = planned / G, extra = planned % G;
			const int64_t lo = g * base + std::min<int64_t>(g, extra), hi = lo + base + (g < extra ? 1 : 0);
			rcs[g] = ssc_generate(devs[g], lo, hi, file_sink, &sinks[g]);
			if (rcs[g]) errs[g] = ssc_last_error();
		};
		std::vector<std::thread> th;
		for (int g = 1; g < G; g++) th.emplace_back(work, g);
		work(0);
		for (auto& t : th) t.join();
		for (int g = 0; g < G; g++) if (rcs[g]) { std::cerr << "Error: " << errs[g] << std::endl; rc = rcs[g]; }
		// ordered concatenation of the shards
		for (int g = 1; g < G && !rc; g++) {
			for (int f = 0; f < (paired ? 2 : 1); f++) {
				const int src = f == 0 ? sinks[g].fd1 : sinks[g].fd2;
				const int dst = f == 0 ? sinks[0].fd1 : sinks[0].fd2;
				close(src);
				const std::string pf = part(f == 0 ? f1 : f2, g);
				int in = open(pf.c_str(), O_RDONLY);
				std::vector<char> buf(8u << 20);
				ssize_t n;
				while (in >= 0 && (n = read(in, buf.data(), buf.size())) > 0) if (write_all(dst, buf.data(), (size_t)n)) { rc = SSC_ERR_SINK; break; }
				if (in >= 0) close(in);
				unlink(pf.c_str());
			}
		}
		close(sinks[0].fd1);
		if (sinks[0].fd2 >= 0) close(sinks[0].fd2);
	}
	for (auto* d : devs) ssc_destroy(d);
	return rc;
}

}  // extern "C"
