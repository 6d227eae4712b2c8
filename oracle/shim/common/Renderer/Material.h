// empty stand-in (rendering is out of scope)
