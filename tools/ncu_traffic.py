"""profiles/r2_traffic.json from the committed ncu tables (profiles/r2_ncu_full_*_b64.csv, written by tools/ncu_table.py
from `ncu --set full` captures of `python tools/profile_layers.py 64 500 fp32`): DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) keyed by the name the library's profile mode gives the launch.
python tools/ncu_traffic.py"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def rows(name):
    with open(os.path.join(P, name)) as f:
        return list(csv.DictReader(f))


def tot(r):
    return float(r["dram_read_B"]) + float(r["dram_write_B"])


k7 = rows("r2_ncu_full_k7_768_and_1x1_b64.csv")
ru = rows("r2_ncu_full_resunit_fused_b64.csv")
up = rows("r2_ncu_full_upsamplers_b64.csv")
table = {
    "conv_gemm_tc cin=768 n=768 phases=1 taps=7 L=4000 bn=256 act=2 res=0": tot(k7[0]),
    "conv_gemm_tc cin=768 n=768 phases=1 taps=1 L=4000 bn=256 act=2 res=1": tot(k7[1]),
    "resunit_fused c=384 dil=1 L=20000": tot(ru[0]),
    "resunit_fused c=192 dil=1 L=80000": tot(ru[3]),
    "resunit_fused c=96 dil=1 L=160000": tot(ru[6]),
    "conv_gemm_tc cin=768 n=1920 phases=5 taps=3 L=4000 bn=192 act=2 res=0": tot(up[0]),
    "conv_gemm_tc cin=384 n=768 phases=4 taps=2 L=20000 bn=192 act=2 res=0": tot(up[1]),
    "conv_gemm_tc cin=192 n=192 phases=2 taps=2 L=80000 bn=96 act=2 res=0": tot(up[2]),
}
out = {"fp32 b64 t500": table,
       "_source": "profiles/r2_ncu_full_*_b64.csv: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                  "--clock-control none, python tools/profile_layers.py 64 500 fp32 (first pass; tools/evidence_run.sh), "
                  "two-term fp32 mode. Algorithmic bytes of the same launches: k7 C=768 1.57e9 (operand planes in + out; the "
                  "excess is conv-halo rows read twice), 1x1 C=768 3.15e9, fused C=384 7.86e9, fused C=192 / C=96 15.7e9."}
with open(os.path.join(P, "r2_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(table, indent=1))
