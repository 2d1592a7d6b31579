"""CPU test of the exact row-span function shared by raster_kernel and fragment_kernel (m2s_span.cuh): compiled for
the host with g++ and checked against the brute-force per-pixel coverage test on random triangles (slivers, ties
on pixel centres, guard-band extremes, R up to 4096)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_span_row_matches_per_pixel_coverage(tmp_path):
    exe = str(tmp_path / "span_harness")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "span_harness.cpp")], check=True)
    r = subprocess.run([exe, "120000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok ")
