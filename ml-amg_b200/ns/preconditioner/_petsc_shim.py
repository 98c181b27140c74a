"""PETSc/firedrake are not part of this build (firedrake glue out of scope).
The PC classes only need the small protocol below; with firedrake installed its PCBase is used."""
try:                                      # pragma: no cover - firedrake is absent in this image
    from firedrake import PCBase          # type: ignore
    from firedrake.petsc import PETSc     # type: ignore
    HAVE_FIREDRAKE = True
except Exception:
    HAVE_FIREDRAKE = False

    class PCBase(object):
        def view(self, pc, viewer=None):
            pass

    PETSc = None


def get_option(prefix, name, default, kind=float):
    """PETSc options DB lookup (MLAMG.py:61-67) with a plain-dict fallback `OPTIONS`."""
    if PETSc is not None:   # pragma: no cover
        opts = PETSc.Options()
        if kind is float:
            return opts.getScalar(prefix + name, default)
        if kind is int:
            return opts.getInt(prefix + name, default)
        if kind is bool:
            return opts.getBool(prefix + name, default)
        return opts.getString(prefix + name, default)
    return kind(OPTIONS.get(prefix + name, default))


OPTIONS = {}
