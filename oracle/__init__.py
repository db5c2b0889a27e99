"""Test infrastructure: CPU oracle for the transport-operator path (see rt_oracle.py)."""
