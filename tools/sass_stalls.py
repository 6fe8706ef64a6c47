"""Static issue-cycle estimate of a SASS region: sums the per-instruction stall fields (bits 105..108 of the 128-bit
encoding, see /opt/skills/guides/B300_MICROARCH.md "Single-warp issue model") between two line patterns.

    cuobjdump -sass lib.so | awk '/Function : /{f=($0 ~ /NAME/)} f' > k.sass
    python tools/sass_stalls.py k.sass [first_addr_hex last_addr_hex]
"""
import re
import sys


def parse(path):
    ins = []
    lines = open(path).read().splitlines()
    i = 0
    pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
    hi = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            h = hi.match(lines[i + 1])
            if h:
                w = int(h.group(1), 16)
                ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(w >> 41) & 0xF, yld=(w >> 45) & 1,
                                wbar=(w >> 46) & 7, rbar=(w >> 49) & 7, wait=(w >> 52) & 0x3F))
                i += 2
                continue
        i += 1
    return ins


if __name__ == "__main__":
    ins = parse(sys.argv[1])
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi_ = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
    sel = [x for x in ins if lo <= x["addr"] <= hi_]
    tot = sum(x["stall"] for x in sel)
    print(f"{len(sel)} instructions, sum of stall fields = {tot} cycles ({tot / max(len(sel), 1):.2f} per instruction)")
    import collections
    byop = collections.Counter()
    cnt = collections.Counter()
    for x in sel:
        op = x["text"].split()[1] if x["text"].startswith("@") else x["text"].split()[0]
        op = op.split(".")[0]
        byop[op] += x["stall"]
        cnt[op] += 1
    for op, c in byop.most_common(14):
        print(f"  {op:10s} n={cnt[op]:4d} stall-sum={c:5d} avg={c / cnt[op]:.2f}")
    if len(sys.argv) > 4:
        for x in sel:
            print(f"{x['addr']:05x} st={x['stall']:2d} y={x['yld']} w={x['wbar']} r={x['rbar']} wm={x['wait']:02x}  {x['text']}")
