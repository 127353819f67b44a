"""Native plug-ins behind the reference's ``dsp.fast`` package name (extra/setup.py:40)."""
