"""Layer library of the IDEAS hot path (module surface of the reference's stylegan2/)."""
