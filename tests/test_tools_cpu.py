"""The measurement tooling behind bench.py's roofline block and the committed profiles, on CPU:
the launch-list summariser reads the committed ncu launch list, and the DRAM-traffic record that bench.py reports is tied to
the kernel sources it was captured from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_reads_the_committed_launch_list():
    csv = os.path.join(ROOT, "profiles", "r2_final_launches.csv")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), csv, "--marker", "adam_polyak_kernel", "--step", "1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [l for l in r.stdout.splitlines() if l and not l.startswith("#")]
    names = " ".join(lines)
    for k in ("gru_step_v2_kernel", "dot_gemm_tc_kernel", "emb_gemm_tc_kernel", "adam_polyak_kernel", "compact_scan_kernel"):
        assert k in names, k
    shares = [float(l.split("%")[0]) for l in lines]
    assert abs(sum(shares) - 100.0) < 0.5
    head = [l for l in r.stdout.splitlines() if l.startswith("#")][-1]
    n = int(head.split()[1])
    assert 250 <= n <= 350, head           # one train step: ~300 launches on the two lanes


def test_traffic_record_is_tied_to_the_sources():
    sys.path.insert(0, ROOT)
    import bench
    rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["gru_step_v2_kernel"]
    assert os.path.exists(os.path.join(ROOT, rec["source"]))
    got, src = bench.measured_traffic("gru_step_v2_kernel", rec["rows"])
    if rec["csrc_hash"] == bench.csrc_hash():
        assert got == rec["dram_bytes_read"] + rec["dram_bytes_write"] and src == rec["source"]
        # the kernel's data flow: a16 images of ae and h, fp32 h_prev in; fp32 h and its image out = 5 x rows x 400 x 4 bytes
        assert 0.9 < got / (5.0 * rec["rows"] * 400 * 4) < 1.15
    else:
        assert got is None and "older kernel sources" in src      # never a number from other code
    assert bench.measured_traffic("gru_step_v2_kernel", rec["rows"] + 1)[0] is None
