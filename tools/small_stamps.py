import os, sys
sys.path.insert(0, "/root/repo")
os.environ["MDB_SMALL_STAMPS"] = "1"
from metabodecon_rust_b200 import Deconvoluter, Spectrum
G = os.path.join("/root/repo", "tests", "golden", "bruker")
sim = Spectrum.read_bruker(os.path.join(G, "sim_01"), 10, 10, (3.34, 3.56))
dec = Deconvoluter()
for _ in range(4):
    dec.deconvolute_spectrum(sim)
