"""Files-to-graph on config-2-sized FASTA text, a few calls -- the command the ncu launch list of the input stage is taken from."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from alga_b200 import synth
from alga_b200.input_reader import FASTA, PinnedText, build_overlap_graph

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(2)
g = synth.make_genome(int(4_600_000 * scale), rng)
m1, m2 = synth.sample_paired_end(g, 150, 50, rng, 0.0)
p1, p2 = PinnedText(synth.fasta_text(m1)), PinnedText(synth.fasta_text(m2))
for _ in range(reps):
    og = build_overlap_graph(p1, p2, FASTA)
    print(og.timing)
