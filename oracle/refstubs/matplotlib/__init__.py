"""Empty stand-in for matplotlib (imported by the reference crypto env, used only for rendering)."""
