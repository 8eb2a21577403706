"""SASS of the hot GEMM instantiations -> profiles/ (cuobjdump -sass of the in-tree object; no GPU needed).

Writes, per requested instantiation of dgemm_dmma_v4_kernel<AK, BKM, BK, HO>, the full listing plus a summary: the
instruction mix of the consumer main loop (the backward-branch loop holding the DMMA.8x8x4 chain) and of the producer
loop (UBLKCP copies), and where the mbarrier instructions (SYNCS) and proxy fences (FENCE.VIEW.ASYNC) sit.

    python scripts/dump_sass.py            # NT BK=16 with the producer-side fence (default) and with the round-1 fold
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "dynaml_b200" / "build" / "gemm.cu.o"
WANT = {
    "sass_gemm_v4_nt_bk16_prodfence_r2.txt": "dgemm_dmma_v4_kernelILb0ELb0ELi16ELi1EE",
    "sass_gemm_v4_nt_bk16_depfold_r2.txt": "dgemm_dmma_v4_kernelILb0ELb0ELi16ELi0EE",
    "sass_gemm_v4_nn_bk32_prodfence_r2.txt": "dgemm_dmma_v4_kernelILb0ELb1ELi32ELi1EE",
}


def functions(text):
    cur, out = None, {}
    for ln in text.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(ln)
    return out


def instructions(lines):
    """[(address, mnemonic, text)] of one function."""
    ins = []
    for ln in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if not m:
            continue
        body = m.group(2).strip()
        body = re.sub(r"^@!?U?P\d+\s+", "", body)
        ins.append((int(m.group(1), 16), body.split()[0], body))
    return ins


def loops(ins):
    """Backward branches: (target, branch address)."""
    out = []
    for addr, mn, body in ins:
        if mn.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", body)
            if m and int(m.group(1), 16) <= addr:
                out.append((int(m.group(1), 16), addr))
    return out


def summarise(name, lines):
    ins = instructions(lines)
    txt = [f"== {name}: {len(ins)} instructions =="]
    for lo, hi in loops(ins):
        body = [(a, m, b) for a, m, b in ins if lo <= a <= hi]
        if len(body) > 300:   # an enclosing region, not a loop of interest
            continue
        mix = Counter(m.split(".")[0] + ("." + m.split(".")[1] if m.startswith(("LDS", "DMMA", "SYNCS", "FENCE", "UBLKCP")) and "." in m else "")
                      for _, m, _ in body)
        if not any(k.startswith(("DMMA", "UBLKCP", "SYNCS")) for k in mix):
            continue
        kind = "consumer main loop" if any(k.startswith("DMMA") for k in mix) else (
            "producer loop" if any(k.startswith("UBLKCP") for k in mix) else "mbarrier wait loop")
        txt.append(f"-- loop 0x{lo:04x}..0x{hi:04x} ({len(body)} instructions): {kind}")
        txt.append("   " + ", ".join(f"{k} x{v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])))
        marks = [(a, b) for a, m, b in body if m.startswith(("SYNCS", "FENCE", "UBLKCP", "ELECT", "WARPSYNC"))]
        last_lds = max((a for a, m, _ in body if m.startswith("LDS")), default=None)
        if last_lds is not None:
            txt.append(f"   last LDS of the loop body at 0x{last_lds:04x}")
        for a, b in marks[:12]:
            txt.append(f"   0x{a:04x}: {b}")
    return "\n".join(txt)


def main():
    if not OBJ.exists():
        sys.exit(f"{OBJ} is missing: run `python -m dynaml_b200.build` first")
    sass = subprocess.run(["cuobjdump", "-sass", str(OBJ)], capture_output=True, text=True, check=True).stdout
    fns = functions(sass)
    out_dir = ROOT / "profiles"
    for fname, key in WANT.items():
        match = [k for k in fns if key in k]
        if not match:
            print(f"no function matching {key}")
            continue
        lines = fns[match[0]]
        head = summarise(match[0], lines)
        # keep address + instruction, drop the encoding words
        slim = []
        for ln in lines:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?;)", ln)
            if m:
                slim.append(f"/*{m.group(1)}*/  {m.group(2)}")
            elif ln.strip().startswith(".") or "Function" in ln:
                slim.append(ln.rstrip())
        (out_dir / fname).write_text(head + "\n\n" + "\n".join(slim) + "\n")
        print(head)


if __name__ == "__main__":
    main()
