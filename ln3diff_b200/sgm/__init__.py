"""Host-side mirror of the reference `sgm` package: only the sampling stack the T23D engine
instantiates (sgm/configs/txt2img-clipl-compat.yaml:12-60)."""
