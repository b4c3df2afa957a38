"""Host-side mirror of the reference `vit` package: the Objaverse AE decoder class
(vit/vit_triplane.py) only."""
