"""Distributions (host-side mirror of `openmcmc.distribution`)."""
