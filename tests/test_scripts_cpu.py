"""Evidence tooling (CPU): the scripts that turn ncu launch lists into the numbers `bench.py` and `profiles/` quote."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _row(i, name, metric, unit, value):
    return f'"{i}","1","python","127.0.0.1","{name}(args)","1","7","(148, 1, 1)","(320, 1, 1)","0","10.0","Command line profiler metrics","{metric}","{unit}","{value}"\n'


def test_conv_traffic_groups_one_step_by_family(tmp_path):
    """`scripts/conv_traffic.py`: one step = stem_conv .. next stem_conv; the pipelined attention kernels (`linattn_kv2_kernel`,
    `linattn_out2_kernel`) belong to the fused-attention family (they were once counted as "other")."""
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
    kernels = ["stem_conv_mma_kernel", "void hd::<unnamed>::conv_gemm_kernel<64, 5, false, 1>", "void hd::<unnamed>::groupnorm_apply_kernel<false, false, false, 8>",
               "void hd::<unnamed>::linattn_kv2_kernel<64>", "hd::<unnamed>::linattn_mix_kernel", "hd::<unnamed>::linattn_out2_kernel",
               "void hd::<unnamed>::linattn_out_kernel<128>", "posterior_step_kernel"]
    body = ""
    i = 0
    for _ in range(3):                       # three steps; the script reads the last complete one
        for k in kernels:
            body += _row(i, k, "gpu__time_duration.sum", "us", "10.0")
            body += _row(i, k, "dram__bytes_read.sum", "Mbyte", "2.0")
            body += _row(i, k, "dram__bytes_write.sum", "Mbyte", "1.0")
            i += 1
    p = tmp_path / "launches.csv"
    p.write_text("==PROF== header line that is not CSV\n" + hdr + body)
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "conv_traffic.py"), str(p), "unet_uncond", "256"],
                         capture_output=True, text=True, check=True).stdout
    fam = json.loads(out)["unet_uncond"]["families"]
    assert json.loads(out)["unet_uncond"]["launches_in_step"] == len(kernels)
    assert fam["conv_gemm"]["launches"] == 1 and fam["groupnorm"]["launches"] == 1
    assert fam["linattn_fused"]["launches"] == 4
    assert fam["other"]["launches"] == 2          # stem + posterior
    assert abs(fam["conv_gemm"]["dram_bytes"] - 3.0e6) < 1 and abs(fam["linattn_fused"]["ncu_ms"] - 0.04) < 1e-9


def test_committed_conv_traffic_matches_the_bench_contract():
    """`bench.py` reads `profiles/conv_traffic.json` for `roofline.traffic`: the families it needs are there, per workload."""
    t = json.loads((ROOT / "profiles" / "conv_traffic.json").read_text())
    for w in ("unet_uncond", "unet_cond"):
        fam = t[w]["families"]
        assert fam["conv_gemm"]["launches"] == 63 and fam["conv_gemm"]["dram_bytes"] > 1e9
        assert t[w]["batch"] == 256 and t[w]["launches_in_step"] == 134
