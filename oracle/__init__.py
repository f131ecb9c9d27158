"""CPU oracle for the DefTet hot path -- TEST INFRASTRUCTURE ONLY.

Restates the reference algorithms (each function cites reference file:line) so that the CUDA kernels in
``deftet_b200/csrc`` can be checked on machines where ``/root/reference`` does not exist.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference`` legs may import, link
or execute anything in this directory; the product package ``deftet_b200`` never does.
"""
