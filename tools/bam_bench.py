"""Throughput of the BAM decoder on the GPU (idl_bam_open) next to the host reader of the stand-in.
  python tools/bam_bench.py [chrom_len_mb] [qual_levels] [level] [reps] [--no-host]
A whole-contig 30x dataset (BASELINE config 1 scaled up) written as BAM; qual_levels 8 gives per-base qualities (a file that compresses about
3:1 like sequencer output; 0 = constant qualities, 12:1).  Reports the stages of idl_bam_open by CUDA events (H2D of the file from pageable
memory, inflate + CRC kernel, record chaining + field extraction), its wall time, and GB/s of compressed input and inflated output.
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from indelope_b200 import cuda, host  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    mb = float(args[0]) if len(args) > 0 else 20.0
    levels = int(args[1]) if len(args) > 1 else 8
    zl = int(args[2]) if len(args) > 2 else 1
    reps = int(args[3]) if len(args) > 3 else 5
    cfg = dict(host.CONFIGS["pr1"]); cfg.update(chrom_len=int(mb * 1e6), n_events=int(200 * mb), qual_levels=levels)
    ds = host.Dataset(**cfg)
    d = tempfile.mkdtemp(prefix="idl_bam_")
    fa, bam = os.path.join(d, "ref.fa"), os.path.join(d, "reads.bam")
    ds.write_fasta(fa); ds.write_bam(bam, level=zl)
    data = open(bam, "rb").read()
    out = {"chrom_len": cfg["chrom_len"], "qual_levels": levels, "zlib_level": zl, "reads": ds.n_reads, "bam_bytes": len(data)}
    if "--no-host" not in sys.argv:
        for th in (1, os.cpu_count() or 1):
            t0 = time.time(); full = host.Dataset.load(fa, bam, threads=th); out["host_load_s_%dthr" % th] = round(time.time() - t0, 3)
            del full
    cuda.Bam(data).close()
    best, walls = None, []
    for _ in range(reps):
        t0 = time.time(); b = cuda.Bam(data); dt = time.time() - t0
        walls.append(round(dt, 4))
        if best is None or dt < best["wall_s"]:
            best = dict(b.info, wall_s=round(dt, 4), n_records=b.n_records)
        b.close()
    out["idl_bam_open"] = best          # the repetition with the shortest wall time
    out["wall_s_all_reps"] = walls
    out["copy_inflate_gbs_in"] = round(best["file_bytes"] / best["ms_inflate"] / 1e6, 2)
    out["copy_inflate_gbs_out"] = round(best["inflated_bytes"] / best["ms_inflate"] / 1e6, 2)
    out["parse_gbs"] = round(best["inflated_bytes"] / best["ms_parse"] / 1e6, 2)
    out["open_wall_gbs_out"] = round(best["inflated_bytes"] / best["wall_s"] / 1e9, 2)
    print(json.dumps(out))
    for f in (fa, fa + ".fai", bam, bam + ".bai"):
        if os.path.exists(f):
            os.remove(f)
    os.rmdir(d)


if __name__ == "__main__":
    main()
