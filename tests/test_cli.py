"""vp8gpu_batch: the batch CLI front end (many files -> host-thread parsing -> one pipelined GPU batch -> files).
Its outputs must be the reference decoder's -yuv / -yuvf / -ppm / -png bytes."""
import hashlib
import subprocess

import pytest

from vp8fix import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def test_batch_cli_outputs_match_reference_digests(lib, golden, tmp_path):
    exe = ROOT / "webp-decoder_b200" / "vp8gpu_batch"
    assert exe.exists(), "run `python __graft_entry__.py build`"
    names = sorted(golden)[::5]
    files = [str(GOLDEN / "webp" / n) for n in names]
    for flag, key, ext in (("-yuv", "yuv", "i420"), ("-yuvf", "yuvf", "i420"), ("-ppm", "ppm", "ppm"), ("-png", "png", "png")):
        out = tmp_path / key
        r = subprocess.run([str(exe), flag, str(out), "--threads", "4", "--chunk", "9", *files], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for n in names:
            data = (out / (n[:-5] + "." + ext)).read_bytes()
            assert hashlib.sha256(data).hexdigest() == golden[n][key], (flag, n)
    # --devices: the files are dealt longest-first to the listed GPUs (the same GPU twice stands in for two here), one
    # pipeline per entry; same bytes
    out = tmp_path / "two"
    r = subprocess.run([str(exe), "-yuvf", str(out), "--devices", "0,0", "--threads", "4", *files], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for n in names:
        assert hashlib.sha256((out / (n[:-5] + ".i420")).read_bytes()).hexdigest() == golden[n]["yuvf"], n
    # a broken input fails that file only, with exit status 1
    bad = tmp_path / "broken.webp"
    bad.write_bytes(b"RIFF\x00\x00\x00\x00WEBPVP8 ")
    r = subprocess.run([str(exe), "-yuvf", str(tmp_path / "mixed"), str(bad), files[0]], capture_output=True, text=True)
    assert r.returncode == 1 and "broken.webp" in r.stderr
    assert (tmp_path / "mixed" / (names[0][:-5] + ".i420")).exists()
