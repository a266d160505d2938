"""Import-path shim: the reference installs its decoders as the `PolarDecoder` package
(PolarDecoder/PolarDecoder/__init__.py); this one forwards to the B200 build."""
