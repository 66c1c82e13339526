"""``fno`` alias for callers that put the repository root (not ``Models/``) on sys.path."""
