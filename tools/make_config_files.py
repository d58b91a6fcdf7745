"""Materialise a BASELINE config as real files: <out>/<cfg>.fa, .fa.fai, .bam, .bam.bai, written by the host stand-in's own
FASTA / BGZF / BAM / BAI writers; `indelope_b200/indelope --min-event-len 5 --min-reads 5 <cfg>.fa <cfg>.bam` then runs the reference's command line on them.
  python tools/make_config_files.py pr1 /tmp/cfg [scale]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from indelope_b200 import host  # noqa: E402


def main():
    name, out = sys.argv[1], sys.argv[2]
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    cfg = dict(host.CONFIGS[name]); cfg["n_events"] = max(1, int(cfg["n_events"] * scale))
    os.makedirs(out, exist_ok=True)
    t0 = time.time(); ds = host.Dataset(**cfg)
    fa, bam = os.path.join(out, name + ".fa"), os.path.join(out, name + ".bam")
    ds.write_fasta(fa); ds.write_bam(bam, level=1)
    print("%s: %d records, %d contigs -> %s (%.1f MB), %s (%.1f MB), %s.bai, %.1f s" % (
        name, ds.n_reads, ds.n_chroms, fa, os.path.getsize(fa) / 1e6, bam, os.path.getsize(bam) / 1e6, bam, time.time() - t0))


if __name__ == "__main__":
    main()
