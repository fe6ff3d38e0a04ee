"""Opcode histogram per kernel from the SASS of the built library (cuobjdump -sass; no GPU needed).
    python tools/sass_opcodes.py aphrodite_engine_b200/libb200decode.so 'marlin_w4a16_tc5_kernel.*bfloat16.*Li0ELi4ELb0ELb0' ...
Prints, for every kernel whose mangled name matches one of the regexes, the instruction count and the opcodes of
interest (tensor-core, TMA / bulk-copy, TMEM, cluster / multimem, mbarrier) followed by the top of the histogram."""
import collections
import re
import subprocess
import sys

INTEREST = ("LDGMC", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOM", "HMMA", "LDSM",
            "MOVM", "SYNCS", "UCGABAR", "CGABAR", "MULTIMEM", "RED", "MEMBAR", "ERRBAR", "F2FP", "LDGSTS", "ELECT", "MAPA",
            "ACQBULK", "BAR", "ATOMG", "UTCCP", "CCTL")


def main():
    lib, pats = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, hist = None, None
    results = []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name is not None:
                results.append((name, hist))
            name = m.group(1) if any(p.search(m.group(1)) for p in pats) else None
            hist = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
        if m:
            hist[m.group(1)] += 1
    if name is not None:
        results.append((name, hist))
    for name, hist in results:
        total = sum(hist.values())
        print(f"== {name}  ({total} instructions)")
        keep = {k: v for k, v in hist.items() if any(k.startswith(i) for i in INTEREST)}
        for k, v in sorted(keep.items(), key=lambda kv: (-kv[1], kv[0])):
            print(f"   {k:<44}{v}")
        print("   top:", ", ".join(f"{k.split('.')[0]}={v}" for k, v in collections.Counter(
            {kk.split('.')[0]: 0 for kk in hist}).items() if False) or ", ".join(
            f"{k}={v}" for k, v in collections.Counter({}).items()) or "", end="")
        base = collections.Counter()
        for k, v in hist.items():
            base[k.split(".")[0]] += v
        print(", ".join(f"{k}={v}" for k, v in base.most_common(12)))


if __name__ == "__main__":
    main()
