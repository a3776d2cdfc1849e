"""Generates tests/golden/depth.K32r3c.s{0,1,2,3}.csv with the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/build_ref.sh) on the sequences of depth_case.py:
    ngskit4b index -i asm.fa -o asm.sfx;  ngskit4b genbioseq -i probes.fa -o probes.seq
    ngskit4b hammings -m0 -K32 -r3 -c -s<S> -i asm.sfx -I probes.seq -o depth.K32r3c.s<S>.csv
usage: python tests/golden/make_depth_golden.py"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import bench  # noqa: E402  (write_fasta)
import depth_case  # noqa: E402
from oracle import hamm_oracle as ho  # noqa: E402

ref = ho.ref_binary(nosleep=True)
asm, probes, _ = depth_case.build()
with tempfile.TemporaryDirectory() as d:
    bench.write_fasta(os.path.join(d, "asm.fa"), asm)
    bench.write_fasta(os.path.join(d, "probes.fa"), probes)
    subprocess.check_call([ref, "index", "-i", "asm.fa", "-o", "asm.sfx", "-r", "asm", "-T8"], cwd=d, stdout=subprocess.DEVNULL)
    subprocess.check_call([ref, "genbioseq", "-i", "probes.fa", "-o", "probes.seq", "-r", "p"], cwd=d, stdout=subprocess.DEVNULL)
    for s in range(4):
        out = "depth.K32r3c.s%d.csv" % s
        subprocess.check_call([ref, "hammings", "-m0", "-K%d" % depth_case.K, "-r%d" % depth_case.R, "-c", "-s%d" % s, "-T8",
                               "-i", "asm.sfx", "-I", "probes.seq", "-o", out], cwd=d, stdout=subprocess.DEVNULL)
        data = open(os.path.join(d, out), "rb").read()
        open(os.path.join(HERE, out), "wb").write(data)
        print(out, len(data), "bytes")
