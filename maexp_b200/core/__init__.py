"""Host-side mirror of the reference's plugin surface (`core/` in Yoshi-0921/MAEXP).

Same factory names, config keys and error behaviour as the reference; the heavy lifting is
done by the CUDA library behind include/maexp_b200.h.
"""
