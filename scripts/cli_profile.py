"""Ad-hoc: where the re-hosted CLI's wall time goes (cProfile, top cumulative entries)."""
import cProfile, pstats, sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dna_factory_b200 import pop_factory
out = tempfile.mkdtemp()
args = ["-s", "10000", "-c", "10000", "-x", sys.argv[1] if len(sys.argv) > 1 else "65536", "-f", "0.01", "-z", "2", "-p",
        "tests/golden/cli_small/deleterious_config.yml", "--outdir", out, "--seed", "4242", "--gpu_select"]
pr = cProfile.Profile(); pr.enable(); pop_factory.main(args); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
