// indelope_b200/csrc/host/indelope_main.cpp -- the `indelope` command line, drop-in for the reference's main module
// (src/indelope.nim:553-608):
//
//     indelope [options] <reference> <BAM>          VCF on stdout
//
// Same options, defaults and output as the reference's docopt block (:556-572).  The host part -- BAM sweep, evidence
// counters, coverage-gap chunking, VCF text -- is the C++ stand-in of libindelope_host.so (the reference keeps it in Nim);
// the BAM is streamed in bounded memory and every region of interest goes through libindelope_cuda.so in batches, two in flight.  There is no CPU path: without a
// CUDA device the program stops with an error.  CRAM input and the undocumented `single-site` debugging mode (:578-586)
// are not provided.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <chrono>
#include <thread>
#include <vector>
#include "indelope_cuda.h"
#include "indelope_host.h"

static const char USAGE[] =
	"indelope 0.0.1 (B200)\n"
	"\n"
	"  Usage: indelope [options] <reference> <BAM>\n"
	"\n"
	"Arguments:\n"
	"\n"
	"  <reference>     reference fasta file.\n"
	"  <BAM>           call variants in this file (coordinate sorted).\n"
	"\n"
	"Options:\n"
	"\n"
	"  -m --min-reads <INT>        minimum number of reads to send for alignment [default: 3]\n"
	"  -c --min-contig-len <INT>   minimum contig length to send for alignment [default: 73]\n"
	"  -e --min-event-len <INT>    minimum size of indel to report [default: 4]\n"
	"  -t --threads <INT>          number of bam decompression threads [default: 1]\n"
	"  -d --device <INT>           CUDA device [default: 0]\n"
	"  -g --gpu-sweep              find the regions of interest on the GPU as well (idl_sweep: the evidence array and the region\n"
	"                              extraction of gen_roi); the BAM is then read whole instead of streamed\n"
	"  -h --help                   show help\n";

struct Lane { idl_batch *batch = nullptr; size_t cap[4] = {0, 0, 0, 0}; };
struct Flight { idlh_rois *rois; uint64_t ticket; };

// INDELOPE_TIMING=1: wall-clock phases on stderr
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int die(const char *what, const std::string &why) { fprintf(stderr, "indelope: %s: %s\n", what, why.c_str()); return 1; }

// --gpu-sweep: BAM records -> arrays -> idl_sweep per target (gen_roi on the GPU) -> batches of regions -> idl_submit.  The host decodes the BAM,
// packs the reads of the regions and writes the VCF; it runs no per-record sweep.
static int main_gpu_sweep(const std::string &fa, const std::string &bam, int min_reads, int min_ctg_len, int min_event_len, int threads, int device)
{
	char err[512] = {0};
	const bool timing = getenv("INDELOPE_TIMING") != nullptr;
	const double t_begin = now_s();
	idl_params P;
	idl_default_params(&P);
	P.min_reads = min_reads; P.min_ctg_len = min_ctg_len; P.min_event_len = min_event_len;
	idl_ctx *ctx = nullptr; int create_rc = IDL_OK;
	std::thread creator([&]() { create_rc = idl_create(device, &P, &ctx); });
	idlh_dataset *ds = idlh_load(fa.c_str(), bam.c_str(), threads, err, sizeof err);
	creator.join();
	if (!ds) { if (ctx) idl_destroy(ctx); return die("input", err); }
	if (create_rc != IDL_OK) { idlh_dataset_free(ds); return die("libindelope_cuda", std::string(idl_strerror(create_rc)) + " (this program has no CPU path; it needs a CUDA device)"); }
	const double t_load = now_s() - t_begin;
	double t_sweep = 0, t_pack = 0, t_wait = 0, t_vcf = 0;
	std::vector<int32_t> rc_, rs_, re_, rn_; std::vector<int64_t> idx_;
	int status = 0;
	for (int32_t c = 0; c < idlh_dataset_n_chroms(ds) && !status; ++c) {
		const std::string name = idlh_dataset_chrom_name(ds, c);
		if (name == "hs37d5" || name.compare(0, 2, "GL") == 0) continue; // skippable, src/indelope.nim:41-42: no record of a decoy contig is cached
		idlh_chrom_reads *cr = idlh_dataset_chrom(ds, c);
		idl_sweep_in in; memset(&in, 0, sizeof in);
		in.chrom_len = cr->chrom_len; in.n_reads = (size_t)cr->n_reads; in.start = cr->start; in.stop = cr->stop; in.flag = cr->flag; in.cigar = cr->cigar; in.cig_off = cr->cig_off;
		idl_sweep_out *so = nullptr;
		const double t1 = now_s();
		const int r = idl_sweep(device, &in, min_reads - 2 > 3 ? min_reads - 2 : 3, min_reads, 600, 0, &so); // gen_roi(b, target, ...), :602
		t_sweep += now_s() - t1;
		if (r != IDL_OK) { status = die("idl_sweep", idl_strerror(r)); idlh_chrom_free(cr); break; }
		for (size_t k = 0; k < so->n_rois; ++k) { rc_.push_back(c); rs_.push_back(so->roi_start[k]); re_.push_back(so->roi_end[k]); rn_.push_back(so->roi_n_reads[k]); }
		for (size_t k = 0; k < so->n_read_idx; ++k) idx_.push_back(so->read_idx[k] + cr->first_read);
		idl_sweep_free(so); idlh_chrom_free(cr);
	}
	idlh_rois *rois = status ? nullptr : idlh_rois_from_regions(ds, (int64_t)rs_.size(), rc_.data(), rs_.data(), re_.data(), rn_.data(), idx_.data());
	if (!status) {
		const idlh_roiset *rs = idlh_rois_view(rois);
		char *h = idlh_vcf_header(rs); fputs(h, stdout); idlh_free(h);
		idlh_vcf *writer = idlh_vcf_new();
		std::vector<Lane> lanes((size_t)(P.n_streams > 0 ? P.n_streams : 1));
		struct Fl { int64_t lo; uint64_t ticket; };
		std::deque<Fl> inflight;
		auto drain = [&]() -> bool {
			const Fl f = inflight.front(); inflight.pop_front();
			const idl_results *res = nullptr;
			double t1 = now_s();
			const int r = idl_wait(ctx, f.ticket, &res);
			t_wait += now_s() - t1; t1 = now_s();
			if (r != IDL_OK) { status = die("idl_wait", std::string(idl_strerror(r)) + " " + idl_last_cuda_error(ctx)); return false; }
			char *txt = idlh_vcf_records(writer, rs, f.lo, &P, res, 0, nullptr);
			fputs(txt, stdout); idlh_free(txt);
			idl_release(ctx, f.ticket);
			t_vcf += now_s() - t1;
			return true;
		};
		size_t nb = 0;
		for (int64_t lo = 0; lo < rs->n_rois && !status; ) {
			int64_t hi = lo, reads = 0;
			while (hi < rs->n_rois && (hi == lo || (reads + rs->roi_n_reads[hi] <= 400000 && hi - lo < 20000))) reads += rs->roi_n_reads[hi++];
			if (inflight.size() >= lanes.size() && !drain()) break;
			Lane &L = lanes[nb % lanes.size()];
			const double t1 = now_s();
			size_t need[4] = {(size_t)(hi - lo), 0, 0, 0};
			idlh_pack_size(rs, lo, hi, &P, &need[1], &need[2], &need[3]);
			bool grow = L.batch == nullptr;
			for (int k = 0; k < 4; ++k) grow |= need[k] > L.cap[k];
			if (grow) {
				if (L.batch) idl_batch_free(ctx, L.batch);
				for (int k = 0; k < 4; ++k) L.cap[k] = need[k] + need[k] / 4 + 64;
				if (idl_batch_alloc(ctx, L.cap[0], L.cap[1], L.cap[2], L.cap[3], &L.batch) != IDL_OK) { status = die("idl_batch_alloc", "out of pinned memory"); break; }
			}
			if (idlh_pack(rs, lo, hi, &P, L.batch) != 0) { status = die("idlh_pack", "the pinned batch is smaller than idlh_pack_size reported (internal error)"); break; }
			uint64_t ticket = 0;
			const int r = idl_submit(ctx, L.batch, &ticket);
			if (r != IDL_OK) { status = die("idl_submit", std::string(idl_strerror(r)) + " " + idl_last_cuda_error(ctx)); break; }
			inflight.push_back({lo, ticket});
			t_pack += now_s() - t1;
			lo = hi; ++nb;
		}
		while (!inflight.empty() && !status) if (!drain()) break;
		for (Lane &L : lanes) if (L.batch) idl_batch_free(ctx, L.batch);
		idlh_vcf_free(writer);
		if (timing)
			fprintf(stderr, "indelope timing (gpu sweep): load %.3f s, idl_sweep %.3f s, pack+submit %.3f s, idl_wait %.3f s, vcf %.3f s, total %.3f s, batches %zu, regions %lld\n",
			        t_load, t_sweep, t_pack, t_wait, t_vcf, now_s() - t_begin, nb, (long long)rs->n_rois);
	}
	if (rois) idlh_rois_free(rois);
	idl_destroy(ctx);
	idlh_dataset_free(ds);
	return status;
}

int main(int argc, char **argv)
{
	int min_reads = 3, min_ctg_len = 73, min_event_len = 4, threads = 1, device = 0;
	bool gpu_sweep = false;
	std::vector<std::string> pos;
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto int_opt = [&](const char *s, const char *l, int &dst) -> int { // 0 no match, 1 consumed, -1 error
			std::string v;
			if (a == s || a == l) { if (i + 1 >= argc) return -1; v = argv[++i]; }
			else if (a.rfind(std::string(l) + "=", 0) == 0) v = a.substr(strlen(l) + 1);
			else if (a.size() > 2 && a.compare(0, 2, s) == 0 && a[1] != '-') v = a.substr(2);
			else return 0;
			char *end = nullptr; const long x = strtol(v.c_str(), &end, 10);
			if (v.empty() || *end) return -1;
			dst = (int)x; return 1;
		};
		if (a == "-h" || a == "--help") { fputs(USAGE, stdout); return 0; }
		if (a == "--version") { puts("indelope 0.0.1"); return 0; }
		if (a == "-g" || a == "--gpu-sweep") { gpu_sweep = true; continue; }
		int rc;
		if ((rc = int_opt("-m", "--min-reads", min_reads)) || (rc = int_opt("-c", "--min-contig-len", min_ctg_len)) ||
		    (rc = int_opt("-e", "--min-event-len", min_event_len)) || (rc = int_opt("-t", "--threads", threads)) || (rc = int_opt("-d", "--device", device))) {
			if (rc < 0) { fputs(USAGE, stderr); return 1; }
			continue;
		}
		if (a.size() > 1 && a[0] == '-') { fputs(USAGE, stderr); return 1; }
		pos.push_back(a);
	}
	if (pos.size() != 2) { fputs(USAGE, stderr); return 1; }

	char err[512] = {0};
	const bool timing = getenv("INDELOPE_TIMING") != nullptr;
	const double t_begin = now_s(); double t_sweep = 0, t_pack = 0, t_wait = 0, t_vcf = 0, t0;
	if (gpu_sweep) return main_gpu_sweep(pos[0], pos[1], min_reads, min_ctg_len, min_event_len, threads, device);
	// gen_roi(b, target, min_read_coverage=min_reads, min_event_support=max(3, min_reads-2)), src/indelope.nim:602; the BAM
	// is swept front to back in bounded memory (idlh_stream_*), a group of regions at a time
	idlh_stream *in = idlh_stream_open(pos[0].c_str(), pos[1].c_str(), threads, min_reads - 2 > 3 ? min_reads - 2 : 3, min_reads, 600, err, sizeof err);
	if (!in) return die("input", err);

	idl_params P;
	idl_default_params(&P);
	P.min_reads = min_reads; P.min_ctg_len = min_ctg_len; P.min_event_len = min_event_len;
	// the CUDA context and the library's workspaces come up on a helper thread while this one already reads the BAM; nothing
	// is printed before the context exists (without a device the program fails with empty stdout)
	idl_ctx *ctx = nullptr;
	int create_rc = IDL_OK, rc = IDL_OK, status = 0; double t_create = 0;
	const double t_open = now_s() - t_begin;
	std::thread creator([&]() { const double t1 = now_s(); create_rc = idl_create(device, &P, &ctx); t_create = now_s() - t1; });
	bool have_ctx = false;
	auto need_ctx = [&]() -> bool {
		if (have_ctx) return true;
		creator.join();
		if (create_rc != IDL_OK) { status = die("libindelope_cuda", std::string(idl_strerror(create_rc)) + " (this program has no CPU path; it needs a CUDA device)"); return false; }
		idlh_rois *t = idlh_stream_targets(in); // echo header % [b.contig_header, "sample"], :599
		char *h = idlh_vcf_header(idlh_rois_view(t)); fputs(h, stdout); idlh_free(h);
		idlh_rois_free(t);
		have_ctx = true;
		return true;
	};
	idlh_vcf *writer = idlh_vcf_new();
	std::vector<Lane> lanes((size_t)(P.n_streams > 0 ? P.n_streams : 1));
	std::deque<Flight> inflight;
	auto drain = [&]() -> bool {
		const Flight f = inflight.front(); inflight.pop_front();
		const idl_results *res = nullptr;
		double t1 = now_s();
		const int r = idl_wait(ctx, f.ticket, &res);
		t_wait += now_s() - t1; t1 = now_s();
		if (r != IDL_OK) { status = die("idl_wait", std::string(idl_strerror(r)) + " " + idl_last_cuda_error(ctx)); idlh_rois_free(f.rois); return false; }
		char *dump = nullptr;
		char *txt = idlh_vcf_records(writer, idlh_rois_view(f.rois), 0, &P, res, 0, &dump);
		fputs(txt, stdout);
		idlh_free(txt); idlh_free(dump);
		idl_release(ctx, f.ticket);
		idlh_rois_free(f.rois);
		t_vcf += now_s() - t1;
		return true;
	};
	// one batch per group of regions, in emission order (the dedup of :604-608 depends on it); while the GPU works on a
	// batch the host reads and sweeps the next stretch of the BAM
	const int64_t target_reads = 400000;
	size_t nb = 0;
	while (!status) {
		err[0] = 0;
		t0 = now_s();
		idlh_rois *grp = idlh_stream_next(in, target_reads, err, sizeof err);
		t_sweep += now_s() - t0;
		if (!grp) { if (err[0]) status = die("input", err); break; }
		const idlh_roiset *rs = idlh_rois_view(grp);
		if (rs->n_rois == 0) { idlh_rois_free(grp); continue; }
		if (!need_ctx()) { idlh_rois_free(grp); break; }
		if (inflight.size() >= lanes.size() && !drain()) { idlh_rois_free(grp); break; }
		Lane &L = lanes[nb % lanes.size()];
		t0 = now_s();
		size_t need[4] = {(size_t)rs->n_rois, 0, 0, 0};
		idlh_pack_size(rs, 0, rs->n_rois, &P, &need[1], &need[2], &need[3]);
		bool grow = L.batch == nullptr;
		for (int k = 0; k < 4; ++k) grow |= need[k] > L.cap[k];
		if (grow) {
			if (L.batch) idl_batch_free(ctx, L.batch);
			for (int k = 0; k < 4; ++k) L.cap[k] = need[k] + need[k] / 4 + 64;
			rc = idl_batch_alloc(ctx, L.cap[0], L.cap[1], L.cap[2], L.cap[3], &L.batch);
			if (rc != IDL_OK) { status = die("idl_batch_alloc", idl_strerror(rc)); idlh_rois_free(grp); break; }
		}
		if (idlh_pack(rs, 0, rs->n_rois, &P, L.batch) != 0) { status = die("idlh_pack", "the pinned batch is smaller than idlh_pack_size reported (internal error)"); idlh_rois_free(grp); break; }
		uint64_t ticket = 0;
		rc = idl_submit(ctx, L.batch, &ticket);
		if (rc != IDL_OK) { status = die("idl_submit", std::string(idl_strerror(rc)) + " " + idl_last_cuda_error(ctx)); idlh_rois_free(grp); break; }
		inflight.push_back({grp, ticket});
		t_pack += now_s() - t0;
		++nb;
	}
	if (!status) need_ctx(); // a BAM without any region still gets its header
	while (!inflight.empty() && !status) if (!drain()) break;
	while (!inflight.empty()) { idlh_rois_free(inflight.front().rois); inflight.pop_front(); }
	for (Lane &L : lanes) if (L.batch && ctx) idl_batch_free(ctx, L.batch);
	idlh_vcf_free(writer);
	if (creator.joinable()) creator.join();
	t0 = now_s();
	if (ctx) idl_destroy(ctx);
	idlh_stream_close(in);
	if (timing)
		fprintf(stderr, "indelope timing: open %.3f s, idl_create %.3f s, read+sweep %.3f s, pack+submit %.3f s, idl_wait %.3f s, vcf %.3f s, teardown %.3f s, total %.3f s, batches %zu\n",
		        t_open, t_create, t_sweep, t_pack, t_wait, t_vcf, now_s() - t0, now_s() - t_begin, nb);
	return status;
}
