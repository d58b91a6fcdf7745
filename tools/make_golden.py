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


FIXTURES = {
    # BASELINE config 1 in small
    "pr1_small": dict(chrom_len=300_000, n_events=60),
    # tandem-repeat rich: most events fall back to the two unbanded alignments per read (src/indelope.nim:312-372)
    "tandem_small": dict(chrom_len=300_000, n_events=60, max_indel=40, tr_fraction=0.8, tr_max_unit=4, seed=31),
}


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, over in FIXTURES.items():
        make(name, over)


def make(name, over):
    cfg = dict(host.CONFIGS["pr1"]); cfg.update(over)
    ds = host.Dataset(**cfg)
    rois = ds.sweep(min_reads=5)
    use_ref = orc.have_ref()
    dump, vcf, cnt = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, use_ref_ksw2=use_ref, dump_level=31)
    d2, v2, _ = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, use_ref_ksw2=False, dump_level=31)
    assert (dump, vcf) == (d2, v2), "lane model and compiled reference DP disagree"
    with open(os.path.join(GOLD, name + ".vcf"), "w") as f:
        f.write(rois.header() + vcf)
    with open(os.path.join(GOLD, name + ".dump"), "w") as f:
        f.write("\n".join(l for l in dump.splitlines() if l[:1] in "RAEV") + "\n")
    print(name, "regions", cnt["regions"], "AL events", cnt["al_events"], "unbanded alignments", cnt["dp_b"], "variants", cnt["variants"], "reference ksw2 used:", use_ref)


if __name__ == "__main__":
    main()
