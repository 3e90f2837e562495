// host_capi.cpp -- C ABI of include/simuscop_host.h and the file sink of the drop-in run.
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

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
	ssc_handle* dev = nullptr;
	if (getenv("SIMUSCOP_PLAN_ONLY")) {
		// host logic only (no GPU): write the SSCPLAN1 dumps of every sample and stop
		const char* prefix = getenv("SIMUSCOP_DUMP_PLAN");
		if (!prefix) { std::cerr << "Error: SIMUSCOP_PLAN_ONLY needs SIMUSCOP_DUMP_PLAN" << std::endl; return 1; }
		for (int s = 0; s < (int)job->job.samples.size(); s++) {
			int64_t a, b;
			int r = job->job.prepare_sample(s, nullptr, std::string(prefix) + "." + std::to_string(s) + ".plan", &a, &b);
			if (r) return r;
		}
		return 0;
	}
	int rc = ssc_create(device, &dev);
	if (rc) { std::cerr << "Error: " << ssc_last_error() << std::endl; return rc; }
	const char* dumpPrefix = getenv("SIMUSCOP_DUMP_PLAN");
	const char* bp = getenv("SIMUSCOP_BATCH_PAIRS");
	if (bp) ssc_set_option(dev, "batch_pairs", atoll(bp));
	sschost::Job& J = job->job;
	for (int s = 0; s < (int)J.samples.size(); s++) {
		const std::string prefix = J.cfg.str["output"] + "/" + J.samples[s].stem;
		FileSink sink;
		const bool paired = J.cfg.paired();
		const std::string f1 = paired ? prefix + "_1.fq" : prefix + ".fq";
		sink.fd1 = open(f1.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
		if (sink.fd1 < 0) sschost::die(-1, "Error: can not open fastq file to save results:\n" + f1);
		if (paired) {
			const std::string f2 = prefix + "_2.fq";
			sink.fd2 = open(f2.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
			if (sink.fd2 < 0) sschost::die(-1, "Error: can not open fastq file to save results:\n" + f2);
		}
		std::string dump;
		if (dumpPrefix) dump = std::string(dumpPrefix) + "." + std::to_string(s) + ".plan";
		int64_t planned = 0, emitted = 0;
		rc = J.prepare_sample(s, dev, dump, &planned, &emitted);
		if (!rc) rc = ssc_generate(dev, 0, planned, file_sink, &sink);
		close(sink.fd1);
		if (sink.fd2 >= 0) close(sink.fd2);
		if (rc) { std::cerr << "Error: " << ssc_last_error() << std::endl; break; }
	}
	ssc_destroy(dev);
	return rc;
}

}  // extern "C"
