"""Framework wrappers around the solvers (counterpart of the reference's ``sunode/wrappers``)."""
