#!/usr/bin/env python3
"""ORACLE SUPPORT (test infrastructure): cut whole function definitions, verbatim, out of a reference source file at BUILD time.

    extract_ref.py OUT SRC 'regex of the definition's first line' [more regexes ...]

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
    out, src, patterns = Path(sys.argv[1]), Path(sys.argv[2]), sys.argv[3:]
    lines = src.read_text(errors="replace").splitlines()
    chunks = []
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
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text("\n".join(chunks))
    return 0


if __name__ == "__main__":
    sys.exit(main())
