"""Byte-compiles the reference's Python layer (src/*.py, src/environments/**) into oracle/_ref/pysrc/src/**/*.pyc.

TEST INFRASTRUCTURE ONLY (part of `make -C oracle ref`).  Nothing is copied: every module is compiled from where it lies under
/root/reference with ``py_compile`` and only the resulting code object is written (a ``.pyc`` image with unchecked-hash
invalidation so no source file is ever looked for, stored as ``<module>.pycode``: the GPU-box snapshot skips ``*.pyc``;
oracle/refstack.py copies the images to ``<module>.pyc`` names in a scratch overlay, the legacy sourceless layout).  The build container and the GPU box run the same CPython,
so the files import there like the compiled ``.so`` modules do.

usage: compile_pysrc.py <reference/src> <out dir>"""
import os
import py_compile
import sys


def main(src: str, out: str) -> int:
    src = os.path.realpath(src)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("cpp", "__pycache__")]
        rel = os.path.relpath(root, src)
        for f in files:
            if not f.endswith(".py"):
                continue
            dst = os.path.join(out, rel, f[:-3] + ".pycode")
            s = os.path.join(root, f)
            if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(s):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(s, cfile=dst, dfile=os.path.join("reference/src", rel, f), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            n += 1
    print(f"oracle/_ref/pysrc: {n} module(s) byte-compiled")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
