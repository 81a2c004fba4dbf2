"""Command line front end of the generation path (``csm-mlx generate`` of the reference, cli/generate.py:72-202)."""
