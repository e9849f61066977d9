"""CPU check of the tie analysis behind the tie-free integer epilogue (fastpcc_b200/csrc/epi_nt.cuh is host + device
code): exhaustive on small shifts, randomised with constructed ties on the shifts of converted models."""
import os.path as osp
import subprocess

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


def test_tie_analysis_host(tmp_path):
    exe = str(tmp_path / 'epi_nt_host')
    subprocess.run(['g++', '-O2', '-o', exe, osp.join(ROOT, 'tests', 'host', 'epi_nt_host.cpp')], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith('ok '), out
