"""Mirror of the reference's ``functions`` package (same module paths, names, signatures)."""
