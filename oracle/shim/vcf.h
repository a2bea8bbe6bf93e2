/* oracle/shim: empty stand-in (nothing from this htslib header is used on the marker path) */
