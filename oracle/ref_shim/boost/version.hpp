// Stand-in for <boost/version.hpp> (test infrastructure; see oracle/ref_shim/README.md). Boost is not in this image: the
// handful of Boost facilities the reference's core / render libraries use are mapped onto the C++17 standard library so that
// the reference's OWN sources compile in place under /root/reference (oracle/Makefile.ref). Nothing here is reference code.
#pragma once
#define BOOST_VERSION 108300
