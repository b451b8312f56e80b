"""Samplers (host-side mirror of `openmcmc.sampler`)."""
