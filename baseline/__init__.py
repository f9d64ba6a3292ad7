"""Harness around the UNMODIFIED reference (installed into baseline/_ref by install_ref.sh)."""
