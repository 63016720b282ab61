"""Race check of the kernels without a GPU: the host emulation under ThreadSanitizer.

`python tests/host_emulation/racecheck.py` rebuilds the emulation of tests/test_device_source_on_host.py
with -fsanitize=thread (EMU_SANITIZE=thread), runs that test file with libtsan preloaded and sorts the
reports.  One OS thread per CUDA thread, __syncthreads / cluster / grid barriers as pthread barriers
(which ThreadSanitizer understands as synchronisation), atomics under a mutex: a shared-memory or
global-memory hazard between two barriers of a kernel - the thing `compute-sanitizer --tool racecheck`
looks for on a device - shows up as a data race here.  A control kernel with a missing __syncthreads()
must be reported, its barriered twin must not.

Reports that are artefacts of the emulation or intended are named and counted, anything else fails:
  static-shared   function-local statics stand in for __shared__; when all blocks of a cooperative /
                  cluster launch run at once they share that copy (every block stores the same
                  thresholds of the current sweep into it)
  worm            k_worm_moves: one thread per experiment flips its own bit of a word with atomicXor
                  while the neighbouring experiments read theirs past L1 - different bits of one word,
                  a race only at word granularity (csrc/moves.cu says so)
  openmp          the oracle's OpenMP loops (libgomp is not instrumented)
Not part of the test suite: ~4 minutes.  Last run of the committed kernels: see DESIGN.md 2.
"""
import glob
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CONTROL = r'''
#include "cuda_on_host.h"
static void k_racy(const int n) {
    uint32_t* sm = emu::dyn_smem;
    sm[threadIdx.x] = threadIdx.x;
    volatile uint32_t v = sm[(threadIdx.x + 1) % n];   // no barrier between the store and the neighbour's load
    (void)v;
}
static void k_barriered(const int n) {
    uint32_t* sm = emu::dyn_smem;
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    volatile uint32_t v = sm[(threadIdx.x + 1) % n];
    (void)v;
}
extern "C" void run(int racy) { emu::launch<int>(racy ? k_racy : k_barriered, dim3(1), dim3(64), 256, 64); }
'''


def reports(pattern):
    out = []
    for path in glob.glob(pattern):
        text = open(path, errors="replace").read()
        out += [r for r in text.split("==================") if "WARNING: ThreadSanitizer" in r]
    return out


def classify(r):
    if "libgomp" in r:
        return "openmp"
    if re.search(r"#0 atomicXor ", r) and "k_worm_moves" in r:
        return "worm"
    m = re.search(r"Location is global '([^']*)'", r)
    if m and re.search(r"k_sweep_stencil_(cluster|coop)<", m.group(1)) and "launch_resident" in r:
        return "static-shared"
    return "UNEXPECTED"


def main():
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        print("libtsan.so not found")
        return 2
    work = tempfile.mkdtemp(prefix="ising_racecheck_")
    env = dict(os.environ, LD_PRELOAD=tsan, EMU_SANITIZE="thread")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")

    # control: the tool must see a missing barrier and must stay silent about the barriered twin
    src, so = os.path.join(work, "control.cpp"), os.path.join(work, "libcontrol.so")
    open(src, "w").write(CONTROL)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-fPIC", "-shared", "-pthread", "-w", "-I", HERE,
                    "-I", cuda_inc, src, "-o", so], check=True)
    for racy in (0, 1):
        log = os.path.join(work, f"control{racy}")
        subprocess.run([sys.executable, "-c", f"import ctypes; ctypes.CDLL({so!r}).run({racy})"],
                       env=dict(env, TSAN_OPTIONS=f"exitcode=0 log_path={log}"), check=True)
        n = len(reports(log + ".*"))
        print(f"control, {'missing' if racy else 'with'} __syncthreads: {n} report(s)")
        if (n > 0) != bool(racy):
            print("the control does not behave: ThreadSanitizer is not looking at the emulation")
            return 2

    log = os.path.join(work, "kernels")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_device_source_on_host.py"), "-q",
                          "-p", "no:cacheprovider", "-k", "not build_variants and not every_kernel"],
                         env=dict(env, TSAN_OPTIONS=f"report_signal_unsafe=0 exitcode=0 log_path={log}"), cwd=ROOT,
                         capture_output=True, text=True)
    print(res.stdout.strip().splitlines()[-1])
    if res.returncode != 0:
        print(res.stdout[-3000:])
        return 1
    counts, unexpected = {}, []
    for r in reports(log + ".*"):
        c = classify(r)
        counts[c] = counts.get(c, 0) + 1
        if c == "UNEXPECTED":
            unexpected.append(r)
    print("reports by kind:", counts or "none")
    for r in unexpected[:5]:
        print(r[:3000])
    return 1 if unexpected else 0


if __name__ == "__main__":
    sys.exit(main())
