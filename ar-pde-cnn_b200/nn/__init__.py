"""Host-side mirror of the reference's nn/ packages (shared implementation; per-project module names
live under ar-pde-cnn_b200/dropin/<project>/nn/)."""
