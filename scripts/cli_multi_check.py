"""Ad-hoc (needs 2 GPUs): the CLI's --gpus 2 output decompresses to the same VCF as --gpus 1."""
import gzip, os, random, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dna_factory_b200 import pop_factory
outs = []
for g in (1, 2):
    d = tempfile.mkdtemp()
    random.seed(7)
    import numpy; numpy.random.seed(7)
    class FD(pop_factory.datetime):
        @classmethod
        def now(cls, tz=None): return cls(2026, 1, 1, 12, 34, 56)
    pop_factory.datetime = FD
    pop_factory.main(["-s", "3000", "-c", "3000", "-x", "3000", "-f", "0.01", "-z", "2", "-p", "tests/golden/cli_small/deleterious_config.yml",
                      "--outdir", d, "--seed", "99", "--gpus", str(g)])
    outs.append(gzip.open(os.path.join(d, "population.vcf.gz"), "rb").read())
print("gpus 1 vs 2: %d bytes, identical: %s" % (len(outs[0]), outs[0] == outs[1]))
