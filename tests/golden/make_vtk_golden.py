"""Record what the UNMODIFIED reference programs write, for the drop-in driver tests.

Compiles /root/reference/{cavity,channel,backwards_step}-01.cpp as their header comments say
(g++ -std=c++17 -O2, plus -ffp-contract=off), runs each in a scratch directory, and stores in
tests/golden/drivers.json: md5 of every vtk_output/*.vtk and *.pvd, and the "Step ..." / banner lines of
stdout with the ANSI colour codes stripped.  Build container only (needs /root/reference):

    python tests/golden/make_vtk_golden.py [--reuse /tmp/refrun] [--cases cavity channel]

The backwards-step program runs ~40 minutes (0.8 s per step, nearly every step at the 10 000-iteration cap); it is
recorded with a wall-clock cut: `--cases backwards_step --cut-s 40` stops the reference after 40 s and keeps the frames
whose "Exported VTK file" line made it to stdout (frames 0, 10, 20, ...), the stdout up to the last of them and the
stderr warnings printed until then.  The PVD collection is written at the end of a run and is not part of that record.
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
    ap.add_argument("--cut-s", type=float, default=0.0, help="stop the reference after this many seconds and record what it wrote until then")
    args = ap.parse_args()
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for case in args.cases:
        d = os.path.join(args.reuse or args.scratch, case)
        if not args.reuse:
            os.makedirs(d, exist_ok=True)
            exe = os.path.join(d, case + ".bin")
            subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", f"/root/reference/{case}-01.cpp", "-o", exe], check=True)
            with open(os.path.join(d, "stdout.log"), "w") as so, open(os.path.join(d, "stderr.log"), "w") as se:
                # stdbuf: the reference's std::cout is block-buffered into a file; line-buffer it so a cut run keeps its log
                cmd = (["stdbuf", "-oL", "-eL"] if args.cut_s else []) + [exe]
                try:
                    subprocess.run(cmd, cwd=d, stdout=so, stderr=se, check=True, timeout=args.cut_s or None)
                except subprocess.TimeoutExpired:
                    pass
        vdir = os.path.join(d, "vtk_output")
        out = ANSI.sub("", open(os.path.join(d, "stdout.log")).read()).splitlines()
        err = ANSI.sub("", open(os.path.join(d, "stderr.log")).read()).splitlines()
        if args.cut_s:  # only frames whose export line was printed are complete; the log ends with the last of them
            done = [l.split(": ", 1)[1] for l in out if l.startswith("Exported VTK file: ")]
            out = out[:max(i for i, l in enumerate(out) if l.startswith("Exported VTK file: ")) + 1]
            files = {f: md5(os.path.join(vdir, f)) for f in done}
        else:
            files = {f: md5(os.path.join(vdir, f)) for f in sorted(os.listdir(vdir))}
        data[case] = {"md5": files, "stdout": out, "stderr_count": len(err), "stderr_first": err[:20 if args.cut_s else 3]}
        if args.cut_s:
            data[case]["cut"] = f"reference stopped after {args.cut_s:.0f} s; {len(files)} complete frames"
        print(case, len(files), "files,", len(out), "stdout lines,", len(err), "stderr lines")
    json.dump(data, open(OUT, "w"), indent=0)


if __name__ == "__main__":
    main()
