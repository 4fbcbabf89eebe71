"""No-op ``colored`` stand-in so the reference's timeline printer imports (oracle, test-only)."""


def attr(_name):
    return ""


def fg(_name):
    return ""


def bg(_name):
    return ""
