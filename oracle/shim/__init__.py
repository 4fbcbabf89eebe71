"""Stand-in modules (``librosa``, ``soundfile``, ``colored``) for the oracle.

Put this directory on ``sys.path`` to let the reference's host code import them by
their real names; or import ``oracle.shim.librosa`` directly.
"""
