"""N1: the reference's train.py, UNMODIFIED, runs on the B200 implementation (north_star: "train.py drops in
unchanged").  The script itself is the only reference file involved -- `models`, `utils`, `dataset`, `stylegan2`
resolve to ideas_b200/compat -- so the test needs a copy of train.py: $IDEAS_REF_TRAIN, /root/reference/train.py
(build container) or baseline/_ref/IDEAS/train.py (staged next to the repo for a GPU run by
scripts/stage_reference.sh; git-ignored).  Without one it is skipped.

18 iterations at batch 2 on a folder of PNG files (the `normal` dataset type: PIL decode, RandomHorizontalFlip,
ToTensor, Normalize) cover the lazy-R1 iteration (16), the evaluation block with the secret-message round trip on
CUDA tensors (train.py:249-305, iteration 17), the checkpoint save (train.py:308-322, iteration 18); a second run
resumes from that checkpoint (train.py:435-442) for two more iterations."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("IDEAS_REF_TRAIN", ""), "/root/reference/train.py",
              os.path.join(ROOT, "baseline", "_ref", "IDEAS", "train.py")]
TRAIN_PY = next((c for c in CANDIDATES if c and os.path.exists(c)), None)


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "ideas_b200.compat.run_train", TRAIN_PY] + args, cwd=cwd, env=env,
                       capture_output=True, text=True, timeout=1500)
    log = r.stdout + "\n--- stderr ---\n" + r.stderr[-6000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):                                   # keep the evidence of a GPU-box run
        with open(os.path.join(out_dir, "train_py_compat.log"), "a") as f:
            f.write("$ train.py " + " ".join(args) + "\n" + log + "\n")
    assert r.returncode == 0, log[-8000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(TRAIN_PY is None, reason="no copy of the reference's train.py available")
def test_unmodified_train_py_runs_evaluates_saves_and_resumes(tmp_path):
    import numpy as np
    from PIL import Image
    data = tmp_path / "images"
    data.mkdir()
    rng = np.random.default_rng(0)
    for i in range(6):
        Image.fromarray(rng.integers(0, 256, (300, 280, 3), dtype=np.uint8)).save(data / f"{i:03d}.png")
    common = ["--exp_name", "n1", "--dataset_path", str(data), "--dataset_type", "normal", "--batch_size", "2",
              "--log_every", "1", "--show_every", "17", "--save_every", "18"]
    out = _run(common + ["--num_iters", "18"], str(tmp_path))
    iters = [int(m) for m in re.findall(r"^\[(\d{7})/0000018\] Total:", out, flags=re.M)]
    assert iters == list(range(1, 19)), iters
    m = re.search(r"\[Testing 0000017/0000018\].*ACC of Msg: ([0-9.]+); L1 loss of tensor: ([0-9.]+)", out)
    assert m, out[-3000:]
    assert 0.0 <= float(m.group(1)) <= 1.0
    exp = tmp_path / "experiments" / "n1"
    assert (exp / "samples" / "0000017.png").exists() and (exp / "checkpoints" / "18.pt").exists()
    totals = [float(v) for v in re.findall(r"Total: ([-0-9.naif]+);", out)]
    assert len(totals) == 18 and all(v == v and abs(v) < 1e4 for v in totals), totals
    # resume (train.py:435-442): start_iter = 18, iterations 19 and 20 run
    out2 = _run(common + ["--num_iters", "20", "--ckpt", "18"], str(tmp_path))
    assert "load model: 18" in out2
    assert [int(m) for m in re.findall(r"^\[(\d{7})/0000020\] Total:", out2, flags=re.M)] == [19, 20]
