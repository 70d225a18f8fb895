"""Record what the UNMODIFIED reference programs write, for the drop-in driver tests.

Compiles /root/reference/{cavity,channel,backwards_step}-01.cpp as their header comments say
(g++ -std=c++17 -O2, plus -ffp-contract=off), runs each in a scratch directory, and stores in
tests/golden/drivers.json: md5 of every vtk_output/*.vtk and *.pvd, and the "Step ..." / banner lines of
stdout with the ANSI colour codes stripped.  Build container only (needs /root/reference):

    python tests/golden/make_vtk_golden.py [--reuse /tmp/refrun] [--cases cavity channel]

The backwards-step program runs ~40 minutes; it is included only when asked for.
"""
import argparse
import hashlib
import json
import os
import re
import subprocess

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "drivers.json")
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reuse", default=None, help="directory with <case>/vtk_output, stdout.log, stderr.log from an earlier run")
    ap.add_argument("--cases", nargs="+", default=["cavity", "channel"])
    ap.add_argument("--scratch", default="/tmp/pm_refrun")
    args = ap.parse_args()
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for case in args.cases:
        d = os.path.join(args.reuse or args.scratch, case)
        if not args.reuse:
            os.makedirs(d, exist_ok=True)
            exe = os.path.join(d, case + ".bin")
            subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", f"/root/reference/{case}-01.cpp", "-o", exe], check=True)
            with open(os.path.join(d, "stdout.log"), "w") as so, open(os.path.join(d, "stderr.log"), "w") as se:
                subprocess.run([exe], cwd=d, stdout=so, stderr=se, check=True)
        vdir = os.path.join(d, "vtk_output")
        files = {f: md5(os.path.join(vdir, f)) for f in sorted(os.listdir(vdir))}
        out = ANSI.sub("", open(os.path.join(d, "stdout.log")).read()).splitlines()
        err = ANSI.sub("", open(os.path.join(d, "stderr.log")).read()).splitlines()
        data[case] = {"md5": files, "stdout": out, "stderr_count": len(err), "stderr_first": err[:3]}
        print(case, len(files), "files,", len(out), "stdout lines,", len(err), "stderr lines")
    json.dump(data, open(OUT, "w"), indent=0)


if __name__ == "__main__":
    main()
