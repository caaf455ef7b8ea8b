"""The gortt command line prints through gort_b200/host/fastfmt.h instead of printf("%f "): the bytes must be the
same for every double (the reference's stdout format, gortt.c:310-327, is part of the drop-in contract)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_fast_formatter_is_byte_identical_to_printf(tmp_path):
    exe = tmp_path / "fastfmt_check"
    subprocess.run(["gcc", "-O2", "-Wall", "-o", str(exe), str(ROOT / "tests" / "c" / "fastfmt_check.c"), "-lm"], check=True)
    p = subprocess.run([str(exe), "3000000", "20261018"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert "0 mismatches" in p.stderr
