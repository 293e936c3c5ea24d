"""TEST STUB: matplotlib is not installed in this image; the reference's plotting modules import it at module level."""
