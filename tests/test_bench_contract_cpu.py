"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the reference's CPU path timed on
the host cores) prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(4, os.cpu_count() or 1)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "CXR images/sec embedded+scored" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_traffic_file_matches_the_committed_launch_list():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    with open(p) as f:
        t = json.load(f)
    assert t["batch"] == 512 and t["conv_launches"] >= 36      # 38 since layer3's identity blocks run chained (42 before)
    assert os.path.exists(os.path.join(ROOT, t["source"]))
    # algorithmic bytes of the conv path are ~237 MB per image; measured DRAM traffic must be in that range
    per_image = t["conv_dram_bytes_per_step"] / 512
    assert 120e6 < per_image < 260e6


def test_traffic_json_is_what_the_committed_ncu_launch_list_says(tmp_path):
    """``roofline.traffic`` comes from ``profiles/traffic.json``; that file must be reproducible from the committed ncu
    launch list it names (one forward = the stem, 37 further convolution launches and the head: three layer1 block kernels,
    four pair-chained layer3 blocks, two chained layer2 blocks and 28 single convolutions)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "traffic.json")) as f:
        committed = json.load(f)
    src = os.path.join(root, committed["source"])
    assert os.path.exists(src), f"{committed['source']} is not committed"
    dst = tmp_path / "copy.csv"
    subprocess.run([sys.executable, os.path.join(root, "tools", "summarise_ncu_launches.py"), src, str(dst),
                    str(committed["batch"])], check=True, capture_output=True)
    with open(tmp_path / "traffic.json") as f:
        again = json.load(f)
    for key in ("conv_launches", "all_launches", "conv_dram_bytes_per_step", "all_dram_bytes_per_step"):
        assert again[key] == committed[key], key
    assert committed["conv_launches"] == 38 and committed["all_launches"] == 39
    # every launch moves at least its algorithmic bytes; the whole step within 10 % of the 237 MB per image of SURVEY 8(d)
    per_image = committed["conv_dram_bytes_per_step"] / committed["batch"]
    assert 150e6 < per_image < 1.1 * 237e6
