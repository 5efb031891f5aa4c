"""Repaired trainer shell around the B200 render hot path."""
