"""To-file run of the re-hosted CLI on a slice of config C4 (100 000 samples), 1 vs N ranks:
    python scripts/cli_multi.py <snps> <gpus> <level> [outdir root, default /dev/shm] [samples per group, default 50000]
Prints write_vcf_snps / total wall per run and the file size.  (Parity of the N-rank file is a test:
tests/test_gpu_cli.py::test_cli_two_gpus_same_vcf_and_index; inflating 26 GB here would cost minutes of box time.)"""
import os, re, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
Z = int(sys.argv[3]) if len(sys.argv) > 3 else 6
base = sys.argv[4] if len(sys.argv) > 4 else "/dev/shm"
N = int(sys.argv[5]) if len(sys.argv) > 5 else 50000
for g in sorted({1, G}):
    env = dict(os.environ)
    out = tempfile.mkdtemp(dir=base)
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "dna_factory_b200.pop_factory", "-s", str(N), "-c", str(N), "-x", str(S), "-f", "0.01",
                        "-z", str(Z), "-p", os.path.join(ROOT, "tests/golden/cli_small/deleterious_config.yml"), "--outdir", out,
                        "--seed", "4242", "--gpu_select", "--gpus", str(g)], capture_output=True, text=True, cwd=ROOT, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        print("gpus=%d FAILED: %s" % (g, (r.stderr or r.stdout)[-600:]))
        shutil.rmtree(out, ignore_errors=True)
        continue
    if os.environ.get("DNAF_TRACE"):
        print(r.stderr[-6000:])
    w = sum(float(x) for x in re.findall(r"Finished write_vcf_snps chunk Elapsed time: ([0-9.]+) seconds", r.stdout))
    size = os.path.getsize(os.path.join(out, "population.vcf.gz"))
    calls = 2 * N * S
    print("gpus=%d -z %d: %d samples x %d SNPs = %.2e calls; write_vcf_snps %.2f s (%.2e calls/s), CLI wall %.2f s; population.vcf.gz %d bytes"
          % (g, Z, 2 * N, S, calls, w, calls / w if w else 0, wall, size), flush=True)
    shutil.rmtree(out, ignore_errors=True)
