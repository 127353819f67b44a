"""Drop-in mirror of the reference's ``src/dsp`` package (same module, class and function names),
backed by libsdrterm_b200.so.  No module here computes on the CPU: every operator and processor
needs the built CUDA library and a GPU, and raises otherwise."""
