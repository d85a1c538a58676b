# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): placeholder filled in below."""
