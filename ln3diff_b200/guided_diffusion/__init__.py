"""Host-side mirror of the reference `guided_diffusion` package: the Gaussian-diffusion sampling
math only (gaussian_diffusion.py, respace.py); training, U-Nets and dist helpers are out of scope."""
