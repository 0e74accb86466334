"""Import shim so the unmodified reference can be imported where biopython is absent."""
