"""Print the SASS instruction stream of the functions of an object file / library whose names match a pattern, without
addresses and encodings -- to compare a kernel before and after a source change that should not change it:
    python tools/sass_fn.py vector_db_id_compression_b200/build/roc_kernels.cu.o 'k_roc_(de|en)code_warp' > a.txt"""
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], re.compile(sys.argv[2])
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    show = False
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            show = bool(pat.search(m.group(1)))
            if show:
                print("== " + m.group(1))
            continue
        if not show:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?;)", line)
        if m:
            print(re.sub(r"\s+", " ", m.group(1)))


if __name__ == "__main__":
    main()
