"""Regenerate tests/golden/* from the CPU oracle (run in the build container; the fixtures are committed).

  python tools/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from indelope_b200 import host  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(GOLD, exist_ok=True)
    cfg = dict(host.CONFIGS["pr1"]); cfg.update(chrom_len=300_000, n_events=60)
    ds = host.Dataset(**cfg)
    rois = ds.sweep(min_reads=5)
    use_ref = orc.have_ref()
    dump, vcf, cnt = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, use_ref_ksw2=use_ref, dump_level=31)
    d2, v2, _ = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, use_ref_ksw2=False, dump_level=31)
    assert (dump, vcf) == (d2, v2), "lane model and compiled reference DP disagree"
    with open(os.path.join(GOLD, "pr1_small.vcf"), "w") as f:
        f.write(rois.header() + vcf)
    with open(os.path.join(GOLD, "pr1_small.dump"), "w") as f:
        f.write("\n".join(l for l in dump.splitlines() if l[:1] in "RAEV") + "\n")
    print("regions", cnt["regions"], "variants", cnt["variants"], "reference ksw2 used:", use_ref)


if __name__ == "__main__":
    main()
