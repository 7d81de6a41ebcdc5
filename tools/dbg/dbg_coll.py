import os, sys, tempfile
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import mbgc_b200
from mbgc_b200.matcher import SeqConfig
from mbgc_b200.synth import synth_collection, write_fasta
from util import read_fasta
d = tempfile.mkdtemp()
gs = synth_collection(4, 300_000, 4321, 1e-2, 1e-4)
paths = []
for i, g in enumerate(gs):
    p = os.path.join(d, f"g{i}.fna"); write_fasta(p, f"syn {i}", g); paths.append(p)
g = mbgc_b200.Matcher(1 << 28, 32, 16, 1, 16, devices=[0])
g.disableSlidingWindow()
c0 = read_fasta(paths[0])[0][1]
g.loadRef(c0, True, True, 0)
def onev(ev):
    print("ev", ev.type, "file", ev.file, "seqLen", ev.seqLen, "nM", ev.nMatches, "kept", ev.kept, "unm", ev.unmatched, "ext", ev.extended, ev.extendedRC, "segs", list(ev.segLen), "nC", ev.nContigs, "dna", ev.dnaLineLen, "refExt", ev.refExtSize, "lock", ev.lockPos, "status", ev.status)
images = [open(p, "rb").read() for p in paths]
out, info = g.compressSequential(images, SeqConfig.for_mode(1, 1 << 20), on_event=onev)
print(info["stats"])
# direct: single-text path
m = g.matchTexts(np.frombuffer(c0, dtype=np.uint8))
print("direct match of G0:", m)
