"""Host-side mirror of the reference `nsr` package for the decode/render half of the hot path:
nsr.triplane.{Triplane, OSGDecoder}, nsr.volumetric_rendering.{renderer, ray_sampler}."""
