#!/usr/bin/env python3
"""ORACLE SUPPORT (test infrastructure): cut whole function definitions, verbatim, out of a reference source file at BUILD time.

    extract_ref.py OUT SRC 'regex of the definition's first line' [more regexes ...]
    extract_ref.py --except OUT SRC 'regex' [...]     the whole file WITHOUT those definitions (a reference translation unit around
                                                      functions that the binding replaces)

Each regex must match exactly one line of SRC; the definition runs from that line to the first following line that is a lone
closing brace at the same indentation.  The output lands under oracle/_ref/gen/ (git-ignored): the reference's text is compiled
from where it lies and is never committed.  Used for the member functions of classes whose translation unit cannot be compiled
as a whole here (src/Frame.cc, src/MapPoint.cc, src/KeyFrame.cc, src/Optimizer.cc need the entire system); see
ref_shim/orbslam_standins.hpp for the class declarations they are compiled against.
"""
import re
import sys
from pathlib import Path


def main() -> int:
    argv = sys.argv[1:]
    invert = bool(argv) and argv[0] == "--except"
    if invert:
        argv = argv[1:]
    out, src, patterns = Path(argv[0]), Path(argv[1]), argv[2:]
    lines = src.read_text(errors="replace").splitlines()
    chunks = []
    cut = []
    for pat in patterns:
        rx = re.compile(pat)
        hits = [i for i, l in enumerate(lines) if rx.search(l)]
        if len(hits) != 1:
            print(f"extract_ref: {pat!r} matches {len(hits)} lines of {src}", file=sys.stderr)
            return 1
        i = hits[0]
        indent = len(lines[i]) - len(lines[i].lstrip())
        closing = " " * indent + "}"
        j = i + 1
        while j < len(lines) and lines[j].rstrip() != closing:
            j += 1
        if j == len(lines):
            print(f"extract_ref: no closing brace for {pat!r} in {src}", file=sys.stderr)
            return 1
        chunks.append(f"// ---- {src}:{i + 1}-{j + 1}\n#line {i + 1} \"{src}\"\n" + "\n".join(lines[i:j + 1]) + "\n")
        cut.append((i, j))
    out.parent.mkdir(parents=True, exist_ok=True)
    if invert:
        keep, pos = [f"#line 1 \"{src}\""], 0
        for i, j in sorted(cut):
            keep += lines[pos:i]
            keep.append(f"// ---- {src}:{i + 1}-{j + 1} removed (replaced by the binding)\n#line {j + 2} \"{src}\"")
            pos = j + 1
        keep += lines[pos:]
        out.write_text("\n".join(keep) + "\n")
    else:
        out.write_text("\n".join(chunks))
    return 0


if __name__ == "__main__":
    sys.exit(main())
