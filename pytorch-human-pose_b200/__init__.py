# The directory name contains a hyphen (the project's name), so it is not importable itself:
# put this directory on sys.path and `import hpdecode` (tests/conftest.py, __graft_entry__.py and
# bench.py do exactly that).
