"""Modules named like the reference's (``models``, ``utils``) so that the reference's
``main.py`` can import this implementation unchanged: see INTEGRATION.md."""
