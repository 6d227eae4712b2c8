// empty: GL interop is not part of the solver path
