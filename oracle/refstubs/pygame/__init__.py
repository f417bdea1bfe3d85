"""Empty stand-in for pygame: the reference imports it at module top level but only
uses it in render paths, which are out of scope (Rendering row)."""


def init():
    pass


def quit():
    pass
