"""convopeq_b200.build: the library is rebuilt when its sources or flags change, not when file times do, and concurrent
importers (one process per GPU) never see a half-written file."""
import os

from convopeq_b200 import build as b


def test_staleness_is_a_content_hash_not_a_file_time():
    b.build()                                   # no-op when up to date
    assert os.path.exists(b.LIB) and b.up_to_date()
    src = os.path.join(b.CSRC, "cpq_eq.cuh")
    st = os.stat(src)
    try:
        os.utime(src, None)                     # newer than the library: still the same content
        assert b.up_to_date()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    assert b.source_hash([]) != b.source_hash(["-DCPQ_EQ_L=16"])      # flags are part of the identity
    assert not b.up_to_date(["-DCPQ_EQ_L=16"])


def test_build_knobs_are_the_sources_compile_time_macros(monkeypatch):
    monkeypatch.setenv("CPQ_EQ_ILP2", "0")      # a macro the kernels guard with #ifndef: a compile flag
    monkeypatch.setenv("CPQ_STAGE_THREADS", "4")   # read with getenv at run time: not a compile flag
    monkeypatch.setenv("CPQ_DITHER_SEGMENTS", "1")
    assert b.build_knobs() == ["-DCPQ_EQ_ILP2=0"]
