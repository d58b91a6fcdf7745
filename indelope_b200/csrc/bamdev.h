// indelope_b200/csrc/bamdev.h -- inside libindelope_cuda.so: building a batch of regions on the device from a resident BAM (bamdev.cu),
// entered by idl_bam_submit / idl_bam_pack (pipeline.cu), which own the lanes' device buffers the batch is written into.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "indelope_cuda.h"

struct BamBatchTotals {            // what the device reports after the records are laid out
	unsigned long long n_seq_bases, n_ref_bases;   // multiples of 64
	unsigned max_trim_len, max_ref_len, max_region_reads, n_small_regions;
	unsigned bad_index;                            // a read index outside the BAM's records, or a region on a target without a reference
	unsigned scratch_set;                          // which of the BAM's scratch sets phase 1 used (phase 2 reads the indices from it)
};

// phase 1: region and read records (quality trim, windows, pool offsets) for the regions given as host arrays; d_region / d_read must
// hold n_regions / n_reads records.  Synchronises the stream once to return the totals.
int bam_batch_records(idl_bam *bam, cudaStream_t st, const idl_params *P, size_t n_regions, const int32_t *roi_chrom, const int32_t *roi_start, const int32_t *roi_end,
                      const int32_t *roi_n_reads, const int64_t *read_idx, size_t n_reads, uint32_t ordinal_base, idl_region *d_region, idl_read *d_read,
                      BamBatchTotals *totals);
// phase 2: the bases of the reads (BAM nibbles -> 2 bits + N plane) and of the reference windows (ASCII -> the same) into the pools
int bam_batch_bases(idl_bam *bam, cudaStream_t st, size_t n_regions, size_t n_reads, idl_region *d_region, const idl_read *d_read, const BamBatchTotals *totals,
                    uint32_t *seq2, uint32_t *seqn, uint32_t *ref2, uint32_t *refn);
int bam_device_of(const idl_bam *bam);
