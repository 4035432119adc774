"""ORACLE - test infrastructure only (see oracle/forward_ref.py header)."""
