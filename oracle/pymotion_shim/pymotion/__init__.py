"""ORACLE-ONLY test infrastructure -- NOT part of the product path.

Minimal restatement of the handful of `upc-pymotion==0.1.10` functions that the
DragPoser reference imports (`/root/reference/python/requirements.txt`).  The
real wheel is not installed, is not in /opt/wheelhouse and there is no network,
so the published semantics are restated here (last row).
Only `oracle/` scripts and `tests/` put this directory on `sys.path`.
"""
