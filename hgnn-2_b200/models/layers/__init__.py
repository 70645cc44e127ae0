"""Mirror of the reference's ``models`` package (same module paths, class names, signatures)."""
