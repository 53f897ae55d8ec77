"""Benchmark scenes written in MOOG's one-task-one-file config style.

Each module exposes `get_config(level) -> dict` of Environment kwargs
(reference: moog/README.md:7-15) and only uses the public `moog` API, so the
same file runs against this repo's `moog` spec package (production / bench)
and against the reference's `moog` (golden-vector generation).
"""
