"""Import shim so that /root/reference imports in a container without mmcv (fixture generation only).
mmcv-full 1.x is a third-party dependency of the reference that is not vendored (README.md:24)."""
