// indelope_b200/csrc/host/dataset.h -- the in-memory dataset the host stand-in sweeps: a reference and coordinate-sorted
// reads with CIGARs, filled either by the synthetic generator (synth_sweep.cpp) or from FASTA + BAM files (bamio.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "indelope_host.h"

struct IdlhEvent { int chrom; int64_t pos; std::string ins; int dlen; bool hom; bool tr; };

struct IdlhReadRec {
	int32_t chrom, start, stop; uint8_t mapq; uint16_t flag; int32_t len; int64_t seq_off; int64_t cig_off; int32_t n_cig; uint64_t order;
};

struct idlh_dataset {
	idlh_synth_params P;
	std::vector<std::string> names;
	std::vector<std::vector<uint8_t>> chroms;   // ASCII, as in the FASTA (case preserved)
	std::vector<IdlhEvent> events;              // planted truth (synthetic data only)
	std::vector<IdlhReadRec> reads;             // coordinate sorted per chromosome, chromosomes in header order
	std::vector<uint8_t> bases, quals;          // ASCII bases (soft clips included), raw phred
	std::vector<uint32_t> cigars;               // BAM encoding len<<4|op (M0 I1 D2 N3 S4 H5 P6 =7 X8)
};
