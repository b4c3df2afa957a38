"""Host-side mirror of the reference `dit` package for the generation hot path (same class
names, constructor signatures and state_dict keys; device work in libln3b200.so)."""
