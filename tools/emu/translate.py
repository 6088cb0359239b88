#!/usr/bin/env python3
"""translate.py — rewrites the CUDA sources of reth_b200/csrc into plain C++ for the CPU emulation (tools/emu).

Only syntax the host compiler cannot parse is touched:
  k<<<grid, block, smem, stream>>>(args)     ->  EMU_LAUNCH((k), grid, block, smem, stream, args)
  extern __shared__ T name[];                ->  T *name = reinterpret_cast<T *>(emu::smem_base());
  __shared__                                 ->  static            (blocks run one at a time)
Everything else (qualifiers, intrinsics, runtime API, CUB) is supplied by cuda_emu.h and tools/emu/include."""
import os
import re
import sys


def match_close(s: str, i: int, open_ch: str, close_ch: str) -> int:
    """index just past the bracket that closes s[i] (which must be open_ch)"""
    assert s[i] == open_ch
    depth = 0
    while i < len(s):
        if s[i] == open_ch:
            depth += 1
        elif s[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def rewrite_launches(s: str) -> str:
    out = []
    pos = 0
    while True:
        k = s.find("<<<", pos)
        if k < 0:
            out.append(s[pos:])
            break
        # callee: identifier, optionally followed by balanced template arguments
        j = k
        if s[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if s[j] == ">":
                    depth += 1
                elif s[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", s[:j])
        assert m, "no callee before <<< at %d" % k
        callee = s[m.start():k]
        e = s.find(">>>", k)
        cfg = s[k + 3:e]
        a0 = e + 3
        while s[a0].isspace():
            a0 += 1
        a1 = match_close(s, a0, "(", ")")
        args = s[a0 + 1:a1 - 1].strip()
        parts = [c.strip() for c in split_top(cfg)]
        while len(parts) < 4:
            parts.append("0" if len(parts) == 2 else "nullptr")
        out.append(s[pos:m.start()])
        out.append("EMU_LAUNCH((%s), %s%s)" % (callee, ", ".join(parts), (", " + args) if args else ""))
        pos = a1
    return "".join(out)


def split_top(s: str):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return parts


def translate(src: str) -> str:
    s = src
    s = re.sub(r"extern\s+__shared__\s+([A-Za-z_0-9]+)\s+([A-Za-z_0-9]+)\s*\[\s*\]\s*;",
               r"\1 *\2 = reinterpret_cast<\1 *>(emu::smem_base());", s)
    s = re.sub(r"\b__shared__\b", "static", s)
    s = rewrite_launches(s)
    # The warp-per-node builder costs ~0.5 ms per node under fibers (every shuffle is 32 context switches); the emulated
    # copy therefore sends only small levels to it.  EMU_WARP_LEVEL_MAX=4096 restores the product threshold.
    s = re.sub(r"(WARP_LEVEL_MAX\s*=\s*)4096", r"\g<1>" + os.environ.get("EMU_WARP_LEVEL_MAX", "192"), s)
    return s


def main():
    src_dir, dst_dir = sys.argv[1], sys.argv[2]
    os.makedirs(dst_dir, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src_dir)):
        if not name.endswith((".cu", ".cuh", ".inl", ".h")):
            continue
        text = open(os.path.join(src_dir, name)).read()
        out = translate(text)
        dst = os.path.join(dst_dir, name[:-3] + ".cpp" if name.endswith(".cu") else name)
        if not os.path.exists(dst) or open(dst).read() != out:
            open(dst, "w").write(out)
        n += out.count("EMU_LAUNCH(")
    print("translated %s -> %s (%d launch sites)" % (src_dir, dst_dir, n))


if __name__ == "__main__":
    main()
